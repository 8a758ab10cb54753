python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/ring: /" | cut -c1-230
GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_noring.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/noring: /" | cut -c1-230
GRAIL_CFG=4 python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/ring cfg4: /" | cut -c1-230
GRAIL_CFG=4 GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_noring.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/noring cfg4: /" | cut -c1-230
python -m pytest tests -m gpu -q > gpurun_out/s16_tests.log 2>&1; tail -5 gpurun_out/s16_tests.log
python scripts/parity_margin.py 2>&1 | tail -8

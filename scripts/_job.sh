python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/fr512: /" | cut -c1-200
for v in fr256 fr1024; do GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_$v.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/$v: /" | cut -c1-200; done
GRAIL_CFG=4 python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/fr512 cfg4: /" | cut -c1-200
GRAIL_CFG=4 GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_fr1024.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/fr1024 cfg4: /" | cut -c1-200
GRAIL_CFG=4 GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_fr256.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/fr256 cfg4: /" | cut -c1-200
python -m pytest tests -m gpu -q > gpurun_out/s18_tests.log 2>&1; tail -4 gpurun_out/s18_tests.log
ncu --set full --clock-control none --import-source on -k regex:k_phase_b -c 1 -o gpurun_out/r2_k_phase_b python scripts/prof_cfg2.py 1 > /dev/null 2>&1

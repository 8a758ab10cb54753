python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/new cfg2: /" | cut -c1-220
GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_lk0.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/old cfg2: /" | cut -c1-220
GRAIL_CFG=4 python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/new cfg4: /" | cut -c1-220
GRAIL_CFG=4 GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_lk0.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/old cfg4: /" | cut -c1-220
GRAIL_CFG=3 python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/new cfg3: /" | cut -c1-220
GRAIL_CFG=3 GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_lk0.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/old cfg3: /" | cut -c1-220
python -m pytest tests -m gpu -q -x 2>&1 | tail -3

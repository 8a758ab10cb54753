ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_phase_b -c 1 python scripts/prof_cfg2.py 1 2>&1 | grep -E "gpu__time_duration" | sed "s/^/new B: /"
python scripts/prof_cfg2.py 4 2>&1 | tail -1 | cut -c1-200
GRAIL_CFG=4 python scripts/prof_cfg2.py 4 2>&1 | tail -1 | cut -c1-200
python -m pytest tests -m gpu -q -x 2>&1 | tail -2

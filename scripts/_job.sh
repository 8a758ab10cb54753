python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --config 3 --steps 10 > gpurun_out/f2_c3_n1.json 2> gpurun_out/f2_c3_n1.err; head -c 330 gpurun_out/f2_c3_n1.json; echo; python scripts/cfg3_time.py 2>&1 | head -2

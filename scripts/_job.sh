python -m pytest tests -m gpu -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_under_ncu.log 2>&1
grep -c k_formant gpurun_out/r2_launches_bench_steps2.csv

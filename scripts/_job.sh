python scripts/phase_chunk_check.py > gpurun_out/s5_phase_check.log 2>&1; grep -E "ALL GOOD|FAIL|config2 mode|config3 mode|config4x4096 mode" gpurun_out/s5_phase_check.log | cut -c1-420
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s5_launches.csv python scripts/prof_cfg2.py 2 > gpurun_out/s5_prof.log 2>&1; tail -1 gpurun_out/s5_prof.log
python -m pytest tests -m gpu -q > gpurun_out/s5_tests.log 2>&1; tail -25 gpurun_out/s5_tests.log

python scripts/phase_chunk_check.py > gpurun_out/s12_phase_check.log 2>&1; grep -E "ALL GOOD|FAIL|Error|error" gpurun_out/s12_phase_check.log | head; grep -E "config2 mode|config4x4096 mode" gpurun_out/s12_phase_check.log | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s12_launches.csv python scripts/prof_cfg2.py 2 > gpurun_out/s12_prof.log 2>&1; tail -1 gpurun_out/s12_prof.log
python -m pytest tests -m gpu -q -x > gpurun_out/s12_tests.log 2>&1; tail -8 gpurun_out/s12_tests.log
for v in occ16; do GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_$v.so python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/$v: /" | cut -c1-300; done

#!/bin/bash
# Round-2 profile set (run on the GPU box through gpurun; outputs under gpurun_out/, summaries are copied to profiles/)
set -x
O=gpurun_out
# launch list of one bench-like step (default plan, config 2): per-launch durations, cold-cache and serialised
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_launches_config2.csv python scripts/prof_cfg2.py 2 > $O/r2_prof.log 2>&1
GRAIL_CFG=4 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_launches_config4_slice.csv python scripts/prof_cfg2.py 2 >> $O/r2_prof.log 2>&1
GRAIL_CFG=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_config3.csv python scripts/prof_cfg2.py 2 >> $O/r2_prof.log 2>&1
# full captures of the kernels of a config-2 step
for k in k_formant k_frequency k_phase_a k_phase_b k_phase_chain k_phase_saw; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o $O/r2_${k} python scripts/prof_cfg2.py 1 > /dev/null 2>&1
done
GRAIL_CFG=4 ncu --set full --clock-control none --import-source on -k regex:k_formant -c 1 -o $O/r2_k_formant_config4 python scripts/prof_cfg2.py 1 > /dev/null 2>&1
ls -la $O/r2_*

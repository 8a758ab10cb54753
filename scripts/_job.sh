O=gpurun_out
python bench.py > $O/f5_c2_n1.json 2> $O/f5_c2_n1.err
python bench.py --config 3 --steps 10 > $O/f5_c3_n1.json 2> $O/f5_c3_n1.err
python bench.py --config 4 --steps 5 > $O/f5_c4_n1.json 2> $O/f5_c4_n1.err
python bench.py --config 5 --steps 5 > $O/f5_c5_n1.json 2> $O/f5_c5_n1.err
bash scripts/make_profiles.sh > $O/r2_make_profiles.log 2>&1
for f in $O/f5_*.json; do echo == $f; head -c 230 $f; echo; done

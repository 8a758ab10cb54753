O=gpurun_out
python bench.py --impl reference > $O/f3_ref_c2.json 2> $O/f3_ref_c2.err
python bench.py > $O/f3_c2_n1.json 2> $O/f3_c2_n1.err
python bench.py --pipeline 0 > $O/f3_c2_n1_inorder.json 2> /dev/null
python bench.py --config 3 --steps 10 > $O/f3_c3_n1.json 2> $O/f3_c3_n1.err
python bench.py --config 4 --steps 5 > $O/f3_c4_n1.json 2> $O/f3_c4_n1.err
python bench.py --config 5 --steps 5 > $O/f3_c5_n1.json 2> $O/f3_c5_n1.err
bash scripts/make_profiles.sh > $O/r2_make_profiles.log 2>&1
python -c "import __graft_entry__ as e; e.smoke()" > $O/f3_smoke.log 2>&1; tail -2 $O/f3_smoke.log
for f in $O/f3_*.json; do echo == $f; head -c 250 $f; echo; done

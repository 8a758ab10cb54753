set -x
O=gpurun_out
python bench.py --impl reference > $O/f1_ref_c2.json 2> $O/f1_ref_c2.err
python bench.py > $O/f1_c2_n1.json 2> $O/f1_c2_n1.err
python bench.py --pipeline 0 > $O/f1_c2_n1_inorder.json 2> /dev/null
python bench.py --config 3 --steps 10 > $O/f1_c3_n1.json 2> $O/f1_c3_n1.err
python bench.py --config 4 --steps 5 > $O/f1_c4_n1.json 2> $O/f1_c4_n1.err
python bench.py --config 5 --steps 5 > $O/f1_c5_n1.json 2> $O/f1_c5_n1.err
for f in $O/f1_*.json; do echo == $f; head -c 400 $f; echo; done

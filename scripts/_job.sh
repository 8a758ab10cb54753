for v in default swap; do
  if [ $v = default ]; then L=""; else L="GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_$v.so"; fi
  env $L python scripts/pipeline_test.py 2>&1 | head -2 | cut -c1-110 | sed "s/^/$v: /"
done

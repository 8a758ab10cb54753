for v in default kt64 kt128; do
  if [ $v = default ]; then L=""; else L="GRAIL_CUDA_LIB=/root/repo/_variants/libgrail_$v.so"; fi
  env $L python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/$v cfg2: /" | cut -c1-200
  env $L GRAIL_CFG=4 python scripts/prof_cfg2.py 4 2>&1 | tail -1 | sed "s/^/$v cfg4: /" | cut -c1-200
done

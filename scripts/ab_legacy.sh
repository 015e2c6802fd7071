for i in 1 2; do
for v in NEW OLD; do
  if [ $v = OLD ]; then export NNS_B200_LIB=$PWD/build/abl/abl_OLDHEAD.so; else unset NNS_B200_LIB; fi
  echo -n "$v: "; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -oE '"ms_per_step": [0-9.]+'
done; done

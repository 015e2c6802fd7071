for i in 1 2; do
for f in "" $(ls build/abl/abl_*.so 2>/dev/null); do
  if [ -n "$f" ]; then export NNS_B200_LIB=$PWD/$f; else unset NNS_B200_LIB; fi
  echo -n "${f:-default}: "; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -oE '"ms_per_step": [0-9.]+'
done; done

for f in "" build/abl/abl_PF8.so build/abl/abl_PF12.so; do
  if [ -n "$f" ]; then export NNS_B200_LIB=$PWD/$f; else unset NNS_B200_LIB; fi
  echo -n "${f:-default PF4}: "; timeout 300 python bench.py --workload slab_cavity16384 --steps 2 --warmup 1 2>&1 | grep -oE '"ms_per_step": [0-9.]+|"kernel_ms": [0-9.]+' | tr '\n' ' '; echo
done

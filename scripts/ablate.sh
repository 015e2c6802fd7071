# Timing ablations of the SOR role of chorin_fd_stream.cu (results are WRONG on purpose; only the phase
# cycle counters matter).  Variant libraries are built into build/abl/  and travel with gpurun.
mkdir -p gpurun_out
for f in build/abl/abl_*.so; do
  name=$(basename $f .so)
  echo "=== $name"
  NNS_STREAM_PROF=1 NNS_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -E "prof.|ms_per_step" | sed 's/"config".*//' | cut -c1-200
done

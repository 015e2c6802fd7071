# Timing ablations of the SOR role of chorin_fd_stream.cu (results are WRONG on purpose; only the phase
# cycle counters matter).  Builds variant libraries into /tmp and runs the bench with NNS_STREAM_PROF=1.
mkdir -p gpurun_out
for v in BASE NOSMEM; do
  name=$(echo $v | tr -d ' -' )
  echo "=== $v"
  NNS_STREAM_PROF=1 NNS_B200_LIB=$PWD/build/abl/abl_$name.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -E "prof.|ms_per_step" | sed 's/"config".*//' | cut -c1-200
done

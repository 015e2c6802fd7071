mkdir -p gpurun_out
export NNS_STREAM_MODE=wave NNS_B200_LIB=$PWD/build/abl/abl_NOSWEEP_FINE.so
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
timeout 200 $CMD > gpurun_out/plain_ns.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k regex:chorin_wave -s 3 -c 1 -f -o gpurun_out/prof_nosweep $CMD > gpurun_out/ncu_ns.log 2>&1
tail -n 2 gpurun_out/ncu_ns.log

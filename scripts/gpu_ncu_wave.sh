mkdir -p gpurun_out
export NNS_STREAM_MODE=wave
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
timeout 200 $CMD > gpurun_out/plain_wave.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k regex:chorin_wave -s 3 -c 1 -f -o gpurun_out/prof_wave_a $CMD > gpurun_out/ncu_wave.log 2>&1
tail -n 3 gpurun_out/ncu_wave.log

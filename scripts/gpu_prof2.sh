mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chorin_stream -s 3 -c 1 -f -o gpurun_out/prof_stream_a $CMD > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu2.log

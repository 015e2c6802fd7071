mkdir -p gpurun_out
CMD="python bench.py --workload slab_cavity4096 --steps 1 --warmup 1"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:slab_sweep -s 400 -c 1 -f -o gpurun_out/prof_slab_a $CMD > gpurun_out/ncu3.log 2>&1
tail -n 3 gpurun_out/ncu3.log

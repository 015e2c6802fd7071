mkdir -p gpurun_out
set -x
./scripts/micro/fp64_pipe > gpurun_out/micro_fp64.txt 2>&1; cat gpurun_out/micro_fp64.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chorin_chip -s 3 -c 1 -f -o gpurun_out/prof_chip_r1c $CMD > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu2.log

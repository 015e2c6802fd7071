mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 3000 gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
nproc; lscpu | grep -E "Model name|Socket|Core|Thread" 
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chorin_chip -s 3 -c 1 -o gpurun_out/prof_chip_r1 $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out

# Round evidence: GPU tests, bench (both arms), ncu launch list and one full capture of the dominant kernel.
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_reference.json 2>&1; cat gpurun_out/bench_r1_reference.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chorin_stream -s 3 -c 1 -f -o gpurun_out/prof_stream_r1 $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log

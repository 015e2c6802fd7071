mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err

mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; cat gpurun_out/bench_stream.json; tail -3 gpurun_out/bench_stream.err

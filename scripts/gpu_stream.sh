mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -5
NNS_STREAM_PROF=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; cut -c1-300 gpurun_out/bench_stream.json; tail -15 gpurun_out/bench_stream.err

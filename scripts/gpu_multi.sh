# Multi-GPU evidence: ensemble bench (weak scaling, no collective) and slab bench (strong scaling) on N GPUs,
# plus the multi-rank GPU tests.
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_ens_n$N.json 2> gpurun_out/bench_ens_n$N.err; cut -c1-420 gpurun_out/bench_ens_n$N.json; tail -2 gpurun_out/bench_ens_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload slab_cavity16384 --steps 2 --warmup 1 > gpurun_out/bench_slab16384_n$N.json 2> gpurun_out/bench_slab.err; cut -c1-420 gpurun_out/bench_slab16384_n$N.json; tail -3 gpurun_out/bench_slab.err
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -3

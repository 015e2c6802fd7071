mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --workload slab_cavity4096 --steps 3 --warmup 1 > gpurun_out/bench_slab4096_n1.json 2> gpurun_out/bench_slab.err; cat gpurun_out/bench_slab4096_n1.json; tail -3 gpurun_out/bench_slab.err
  python bench.py --workload slab_cavity16384 --steps 2 --warmup 1 > gpurun_out/bench_slab16384_n1.json 2> gpurun_out/bench_slab.err; cat gpurun_out/bench_slab16384_n1.json; tail -3 gpurun_out/bench_slab.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload slab_cavity16384 --steps 2 --warmup 1 > gpurun_out/bench_slab16384_n$N.json 2> gpurun_out/bench_slab.err; cat gpurun_out/bench_slab16384_n$N.json; tail -3 gpurun_out/bench_slab.err
fi

N=${1:-1}
for tr in 128 64 48 32; do
  echo -n "N=$N TR=$tr: "
  if [ "$N" = "1" ]; then
    NNS_SLAB_TR=$tr timeout 300 python bench.py --workload slab_cavity16384 --steps 2 --warmup 1 2>&1 | grep -oE '"ms_per_step": [0-9.]+|"kernel_ms": [0-9.]+|"ticks_per_step": [0-9]+' | tr '\n' ' '
  else
    NNS_SLAB_TR=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload slab_cavity16384 --steps 2 --warmup 1 2>&1 | grep -oE '"ms_per_step": [0-9.]+|"kernel_ms": [0-9.]+|"ticks_per_step": [0-9]+' | tr '\n' ' '
  fi
  echo
done

mkdir -p gpurun_out
timeout 90 python scripts/gpu_wave_check.py 333 3 2>&1 | tail -6
NNS_STREAM_PROF=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -E "prof.|ms_per_step" | sed "s/\"config\".*//" | cut -c1-250 | grep -v "      0 cycles"
bash scripts/ablate.sh | grep -v "      0 cycles"

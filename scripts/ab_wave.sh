echo -n "legacy default: "; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | grep -oE '"ms_per_step": [0-9.]+'
export NNS_STREAM_MODE=wave
bash scripts/ab_var.sh

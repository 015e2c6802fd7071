for pf in 2 8; do echo "== PF $pf"; NNS_B200_LIB=$PWD/build/abl/pf$pf.so python bench.py --workload slab_cavity4096 --steps 2 --warmup 1 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'])"; done

"""End-to-end step (nns_chorin_fd_step_host, pinned host buffers) for several chunk counts (NNS_STEP_HOST_CHUNKS)."""
import os, sys, time, subprocess
for n in sys.argv[1:] or ["8", "16", "32"]:
    env = dict(os.environ, NNS_STEP_HOST_CHUNKS=n)
    r = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-extras", "--no-cpu-baseline", "--e2e-steps", "6"],
                       capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import json
    d = json.loads(r.stdout.strip().splitlines()[-1])
    cells = 4096 * 128 * 128
    print("chunks", n, "e2e %.3e cell-updates/s = %.1f ms/step" % (d["e2e"]["value"], cells / d["e2e"]["value"] * 1e3))

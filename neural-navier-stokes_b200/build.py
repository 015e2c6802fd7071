"""Build libnns_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python neural-navier-stokes_b200/build.py [--force] [--verbose]
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("NNS_B200_LIB") or os.path.join(HERE, "libnns_b200.so")   # override: experiments only

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--cudart", "static", "-ldl"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("NNS_EXTRA_NVCC_FLAGS", "").split() + \
        (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc wrapper without libstdc++ specs for nvcc's host pass
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd,
                       capture_output=True, text=True, env=env)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""In-tree build of libvcs_b200.so (nvcc, sm_100a only).  Run: python -m vcs_h264_b200.build"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "vcs_b200.cu")
OUT = os.path.join(HERE, "libvcs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the float64 DCT path must not be re-contracted: explicit __fma_rn only (dct_stage.cuh)
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
    "-I", os.path.join(HERE, "..", "include"),
]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(HERE, "..", "include", "vcs_b200.h")]


def stale():
    return not os.path.exists(OUT) or any(os.path.getmtime(OUT) < os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libvcs_b200.so")
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(OUT)

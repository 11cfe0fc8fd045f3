"""Builds nlp_b200/liblda_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "csrc", f) for f in ("lda_b200.cu", "index_b200.cu")]
DEPS = SRCS + [os.path.join(HERE, "csrc", f) for f in ("kernels.cuh", "pcg.cuh", "nccl_dl.h", "index_kernels.cuh")] + [
    os.path.join(HERE, "..", "include", f) for f in ("lda_b200.h", "lda_b200_index.h")]
OUT = os.path.join(HERE, "liblda_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC,-pthread", "-cudart", "static", "--expt-relaxed-constexpr",
]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRCS + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building liblda_b200.so")
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(OUT)

"""Build libsnb.so (the sm_100a CUDA library + C ABI) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the repo snapshot; nvcc cross-compiles
for sm_100a without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libsnb.so")
SOURCES = ["snb_host.cu", "k1_sample_encode.cu", "k2_gemm.cu", "k2_chain.cu", "k2_mlp.cu", "k3_composite.cu", "k_aux.cu", "k_fp32.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]
if os.environ.get("SNB_EXPERIMENTS"):   # tools/exp_chain.py: deliberately wrong kernels that locate a bottleneck (never shipped)
    FLAGS = FLAGS + ["-DSNB_EXPERIMENTS"]


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "snb.h")]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the sources changed since the last build; return the library path."""
    stamp = LIB + ".hash"
    want = _source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libsnb.so must be built where the CUDA toolkit is installed")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + FLAGS + ["-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    open(stamp, "w").write(want)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

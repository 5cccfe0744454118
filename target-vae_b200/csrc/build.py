"""Builds libtvae_b200.so in-tree with nvcc for sm_100a (no torch headers, no libcuda link).

    python target-vae_b200/csrc/build.py [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tvae_b200", "libtvae_b200.so")
SOURCES = ["tvae_api.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _digest():
    h = hashlib.sha256()
    for root in (HERE, os.path.join(os.path.dirname(os.path.dirname(HERE)), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    return h.hexdigest()


def build(force=False, verbose=False, probe=False):
    """probe=True (or TVAE_PROBE=1 on the command line): the development build with the clock64 probes of tc_gemm2.cuh,
    written next to the product library as libtvae_b200_probe.so (tools/probe_pair.py loads it through TVAE_LIB)."""
    out = OUT.replace(".so", "_probe.so") if probe else OUT
    stamp = out + ".sha256"
    dig = _digest()
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")] + (["-DTVAE_PROBE"] if probe else [])
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(HERE, s) for s in SOURCES] + ["-o", out]
    subprocess.check_call(cmd, cwd=HERE)
    with open(stamp, "w") as f:
        f.write(dig)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, probe=os.environ.get("TVAE_PROBE") == "1"))

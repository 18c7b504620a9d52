"""Build libcsn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m cerebralsignalnetworks_b200.build [--force] [--verbose]

The .so lands next to this file so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
LIB_PATH = os.path.join(HERE, "libcsn_b200.so")
SOURCES = ["core.cu", "sosfilt.cu", "dino_loss.cu", "optim_elementwise.cu", "optim_fused.cu", "gemm_f32.cu", "lstm_f32.cu",
           "lstm_api.cu", "gemm_tc.cu", "lstm_tc.cu", "dbg_umma.cu", "lstm_tc_large.cu", "lstm_cluster.cu", "dp_peer.cu", "retrieval.cu", "distill_losses.cu", "dataset.cu", "head_fused.cu", "batchnorm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "csn_b200.h")]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and (f.endswith(".cu") or f.endswith(".cuh") or f.endswith(".h")):
            with open(p, "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ_DIR, "digest.txt")
    dig = _digest()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp) and open(stamp).read() == dig:
        return LIB_PATH
    if not os.path.isfile(NVCC):
        raise RuntimeError("nvcc not found at %s and no up-to-date %s present" % (NVCC, LIB_PATH))
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stderr[-6000:]))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print("built", path)

"""Builds libsvanon_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m streamvoiceanon_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libsvanon_b200.so"
SOURCES = ["gemm.cu", "gemm_pipe.cu", "gemm_tc.cu", "gemm_pair.cu", "conv_small.cu", "kernels_misc.cu", "attn.cu", "ar_decode.cu", "ar_decode_staged.cu", "ar_batch.cu", "ar_attn_tma.cu", "engine.cu", "voc_stream.cu", "api.cu", "batch.cu", "enc_stream.cu", "speaker.cu", "chain.cu", "enc_chain.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "177"]


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*.cu*")) + list(CSRC.glob("*.hpp")) +
                 [PKG.parent / "include" / "svanon.h"])
    return obj.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False) -> Path:
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    jobs = []
    for s in SOURCES:
        obj = objdir / (s + ".o")
        if force or _stale(obj, CSRC / s):
            jobs.append([NVCC, *FLAGS, "-c", str(CSRC / s), "-o", str(obj)])
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr)
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(run, jobs))
    if jobs or not OUT.exists():
        run([NVCC, "-shared", "-o", str(OUT), *[str(objdir / (s + ".o")) for s in SOURCES],
             "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

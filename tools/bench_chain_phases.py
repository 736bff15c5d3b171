"""Phase times of the chain kernel on WARM data: `repeat` x [GEMM phase, element-wise phase] of one launch over the same
operands (svanon_debug_chain_gemm), printed by the kernel's own profile (SVANON_CHAIN_PROF=<n-th launch>).  Compared with the
in-situ profile of the streaming loop (weights, biases and descriptors cold in L2 every chunk) it separates what a phase costs
by itself from what the cold operands add.

    SVANON_CHAIN_PROF=3 python tools/bench_chain_phases.py M N K [repeat]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import _lib  # noqa: E402
from streamvoiceanon_b200.engine import Engine, ptr  # noqa: E402


def main():
    M, N, K = (int(x) for x in sys.argv[1:4])
    repeat = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    lib = _lib.load()
    eng = Engine.get(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    for _ in range(4):
        _lib.check(lib.svanon_debug_chain_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 1, repeat, None))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()

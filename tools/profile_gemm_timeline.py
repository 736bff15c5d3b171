"""In-kernel timeline of one tcgen05 GEMM CTA (tuning build with -DSVANON_TC_PROF, SVANON_LIB=...): clock64 marks of the
first producer thread and the MMA thread of CTA (0,0,0), printed by the library after each svanon_debug_gemm."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import _lib  # noqa: E402
from streamvoiceanon_b200.engine import Engine, ptr  # noqa: E402

eng, lib = Engine.get(0), _lib.load()
for (M, N, K) in ((128, 1536, 512), (128, 512, 1536), (512, 1536, 384), (512, 384, 1536), (16384, 2048, 512)):
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for i in range(3):
        if i == 2:
            print(f"M={M} N={N} K={K}", file=sys.stderr, flush=True)
        _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 0, None))

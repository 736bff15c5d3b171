"""A few launches of the tensor-core GEMM on one large-M shape (many-stream encoder MLP) for `ncu --set full`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import _lib  # noqa: E402
from streamvoiceanon_b200.engine import Engine, ptr  # noqa: E402

M, N, K = (int(x) for x in (sys.argv[1:4] or (16384, 2048, 512)))
eng = Engine.get(0)
lib = _lib.load()
A = torch.randn(M, K, device="cuda")
W = torch.randn(N, K, device="cuda")
b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
for _ in range(4):
    _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 1, None))
torch.cuda.synchronize()
print("done")

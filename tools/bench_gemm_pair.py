"""Pair GEMM kernel (csrc/gemm_pair.cu) against the single-CTA tensor-core kernel (csrc/gemm_tc.cu): bitwise comparison with
a description of where the results differ, fp64 error, and CUDA-event timings of both on the many-stream shapes.

    python tools/bench_gemm_pair.py [check] [time]
"""
import os
import sys
from pathlib import Path

os.environ.setdefault("SVANON_GEMM_PAIR_MIN_M", "256")        # the check runs small shapes through the pair kernel too
os.environ.setdefault("SVANON_GEMM_PAIR_MIN_TILES", "1")

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import _lib  # noqa: E402
from streamvoiceanon_b200.engine import Engine, ptr  # noqa: E402

eng = Engine.get(0)
lib = _lib.load()
WHAT = [a for a in sys.argv[1:]] or ["check", "time"]


def run(pair_mode, A, W, b, M, N, K, act=0):
    out = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.svanon_set_gemm_pair(pair_mode))
    n0 = lib.svanon_gemm_pair_launches()
    _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, act, None))
    torch.cuda.synchronize()
    took = lib.svanon_gemm_pair_launches() - n0
    assert took == (1 if pair_mode else 0), f"pair kernel launches: {took} in mode {pair_mode}"
    return out


def describe(name, out, ref):
    bad = (out != ref) | out.isnan()
    nb = int(bad.sum())
    if nb == 0:
        print(f"    {name}: bitwise equal")
        return True
    rows = bad.any(dim=1).nonzero().flatten()
    cols = bad.any(dim=0).nonzero().flatten()
    d = (out.double() - ref.double()).abs()
    d[out.isnan()] = float("inf")
    print(f"    {name}: {nb} of {bad.numel()} differ ({int(out.isnan().sum())} nan); rows {int(rows.min())}..{int(rows.max())} "
          f"({rows.numel()} rows), cols {int(cols.min())}..{int(cols.max())} ({cols.numel()} cols); max |diff| {float(d[~out.isnan()].max()) if (~out.isnan()).any() else -1:.3e}")
    rb = torch.zeros(8, dtype=torch.long)
    for i in range(8):
        rb[i] = int(bad[(torch.arange(out.shape[0], device="cuda") % 256) // 32 == i].sum())
    cb = torch.zeros(8, dtype=torch.long)
    for i in range(8):
        cb[i] = int(bad[:, (torch.arange(out.shape[1], device="cuda") % 256) // 32 == i].sum())
    print(f"      by row%256//32: {rb.tolist()}   by col%256//32: {cb.tolist()}")
    return False


_lib.check(lib.svanon_debug_gemm_weights_static(2))     # W treated like an engine weight (converted copies cached per pointer)
_lib.check(lib.svanon_set_gemm_mode(2))
KEEP = []                                                # ... so no W address may be reused by another matrix
if "check" in WHAT:
    ok_all = True
    for (M, N, K, act) in [(4096, 256, 32, 0), (4096, 256, 64, 0), (4096, 128, 64, 0), (4096, 512, 512, 0), (16384, 2048, 512, 1),
                           (16384, 512, 2048, 0), (20992, 384, 1536, 0), (20992, 1536, 384, 1), (16500, 256, 96, 0), (5000, 640, 160, 0)]:
        g = torch.Generator(device="cuda").manual_seed(M + N + K)
        A = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        b = torch.randn(N, device="cuda", generator=g)
        KEEP.append(W)
        print(f"M={M} N={N} K={K} act={act}")
        ref = run(0, A, W, b, M, N, K, act)
        r64 = A.double() @ W.double().T + b.double()
        if act == 1:
            r64 = torch.nn.functional.gelu(r64)
        print(f"    single-CTA kernel vs fp64: {float((ref.double() - r64).abs().max()):.3e}")
        for mode in (1, 2):
            out = run(mode, A, W, b, M, N, K, act)
            ok_all &= describe(f"pair mode {mode}", out, ref)
            print(f"    pair mode {mode} vs fp64: {float((out.double() - r64).abs().max()):.3e}")
    print("CHECK", "OK" if ok_all else "FAILED")
if "few" in WHAT:          # for ncu: two launches of each kernel on three shapes
    for (M, N, K) in [(16384, 2048, 512), (16384, 512, 2048), (20992, 384, 1536)]:
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda")
        KEEP.append(W)
        b = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        for mode in (0, 1):
            _lib.check(lib.svanon_set_gemm_pair(mode))
            for i in range(3):
                lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 1, None)
        torch.cuda.synchronize()
if "time" in WHAT:
    for (M, N, K) in [(16384, 2048, 512), (16384, 512, 2048), (16384, 1536, 512), (16384, 512, 512), (16384, 512, 1536),
                      (20992, 1536, 384), (20992, 384, 1536), (20992, 512, 128), (20992, 128, 512), (20992, 1024, 256),
                      (20992, 256, 1024), (5248, 2048, 512), (5248, 512, 2048), (4096, 1536, 384)]:
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda")
        KEEP.append(W)
        b = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        res = {}
        for mode in (0, 1):
            _lib.check(lib.svanon_set_gemm_pair(mode))
            for i in range(5):
                lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 0, None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 50
            e0.record()
            for i in range(n):
                lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 0, None)
            e1.record()
            torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / n * 1e3
        gf = 2 * M * N * K / 1e9
        print(f"M={M:5d} N={N:4d} K={K:4d}: single-CTA {res[0]:6.1f} us ({gf / res[0] * 1e3:6.1f} TFLOP/s fp32-equiv)   "
              f"pair (incl. lo split pass) {res[1]:6.1f} us ({gf / res[1] * 1e3:6.1f})   x{res[0] / res[1]:.2f}")
_lib.check(lib.svanon_set_gemm_pair(-1))
_lib.check(lib.svanon_debug_gemm_weights_static(0))

"""Micro-benchmark of the GEMM back ends on encoder-window shapes: warm (same weights every call, L2-resident) vs
cold (cycling through 64 different weight matrices > L2) timings, CUDA events.

    python tools/bench_gemm.py [modes ...] [--static] [--half]
      --static   weights treated like engine weights: B operand by TMA from pre-tiled copies (kept across calls)
      --half     perf mode (fp16 single-pass)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import _lib  # noqa: E402
from streamvoiceanon_b200.engine import Engine, ptr  # noqa: E402

eng = Engine.get(0)
lib = _lib.load()
SHAPES = [(512, 1536, 384), (512, 384, 1536), (512, 2048, 512), (512, 512, 2048), (128, 1536, 512), (128, 512, 1536),
          (512, 2050, 2048), (256, 128, 1408), (32, 256, 2816), (16, 2304, 768), (16, 768, 2304), (64, 2304, 768),
          (256, 2304, 768), (4096, 1536, 384), (16384, 2048, 512), (16384, 512, 2048), (8192, 128, 1408)]
STATIC, HALF = "--static" in sys.argv, "--half" in sys.argv
MODES = [int(x) for x in sys.argv[1:] if not x.startswith("--")] or [1, 2]
_lib.check(lib.svanon_debug_gemm_weights_static(2 if STATIC else 0))
_lib.check(lib.svanon_set_precision(1 if HALF else 0))
for mode in MODES:
    _lib.check(lib.svanon_set_gemm_mode(mode))
    for (M, N, K) in SHAPES:
        A = torch.randn(M, K, device="cuda")
        nW = max(2, min(64, int(400e6 / (N * K * 4))))
        Ws = [torch.randn(N, K, device="cuda") for _ in range(nW)]
        b = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        res = {}
        for name, pick in (("warm", lambda i: Ws[0]), ("cold", lambda i: Ws[i % nW])):
            for i in range(5):
                lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(pick(i)), ptr(b), ptr(out), M, N, K, 0, None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 100
            e0.record()
            for i in range(n):
                lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(pick(i)), ptr(b), ptr(out), M, N, K, 0, None)
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / n * 1e3
        gf = 2 * M * N * K / 1e9
        print(f"mode {mode} M={M:4d} N={N:4d} K={K:4d}: warm {res['warm']:6.1f} us ({gf / res['warm'] * 1e3:6.1f} TFLOP/s)  "
              f"cold {res['cold']:6.1f} us ({gf / res['cold'] * 1e3:6.1f} TFLOP/s)")
_lib.check(lib.svanon_set_gemm_mode(2))
_lib.check(lib.svanon_debug_gemm_weights_static(0))
_lib.check(lib.svanon_set_precision(0))

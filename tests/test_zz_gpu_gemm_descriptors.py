"""The GEMM descriptors that only the speaker encoders issue (csrc/speaker.hpp), one at a time through
svanon_debug_gemm_taps against an fp64 product on the GPU -- the bisection tool for tests/test_zz_gpu_speaker.py:
row-offset taps of BOTH signs (non-causal dilated convs), a thin N = 32 output written into a column slice of a wide
concat buffer, stride-2 overlapping rows (the TDNN), an A operand that is itself a column slice (lda > K)."""
import ctypes as C

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]

# name, M, N, K, lda, a_row_step, taps (row offsets), ldc, c_col0, a_col0
CASES = [
    ("cam_local_conv_d2", 249, 32, 128, 128, 1, (-2, 0, 2), 1024, 352, 0),
    ("cam_local_conv_d1", 249, 32, 128, 128, 1, (-1, 0, 1), 512, 128, 0),
    ("cam_local_conv_tiny", 2, 32, 128, 128, 1, (-2, 0, 2), 512, 480, 0),
    ("tdnn_stride2_k5", 249, 128, 1600, 320, 2, (0,), 512, 0, 0),
    ("ecapa_layer1_k5", 251, 512, 640, 128, 1, (0,), 512, 0, 0),
    ("res2_conv_d4", 251, 64, 64, 64, 1, (-4, 0, 4), 512, 192, 0),
    ("se_block_in_from_cat_slice", 251, 512, 512, 1536, 1, (0,), 512, 0, 512),
    ("dense_bottleneck_cin992", 249, 128, 992, 992, 1, (0,), 128, 0, 0),
    ("perceiver_kv", 283, 1024, 128, 128, 1, (0,), 1024, 0, 0),
    ("ecapa_cat_conv", 251, 1536, 1536, 1536, 1, (0,), 1536, 0, 0),
]


@pytest.mark.parametrize("mode", [2, 1])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_speaker_gemm_descriptor(case, mode):
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    name, M, N, K, lda, step, taps, ldc, c0, a_col0 = case
    eng, lib = Engine.get(0), _lib.load()
    g = torch.Generator(device="cuda").manual_seed(len(name))
    margin = max(0, -min(taps))
    rows = margin + (M - 1) * step + max(taps) + 1 + (K + lda - 1) // lda        # every tap of every row stays inside
    A = torch.randn(rows, lda, device="cuda", generator=g)
    W = torch.randn(len(taps), N, K, device="cuda", generator=g) / (K * len(taps)) ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    Cbuf = torch.full((M, ldc), 7.0, device="cuda")
    offs = (C.c_int * len(taps))(*taps)
    a_base = A.reshape(-1)[a_col0:]                                              # a column slice of a wider buffer
    _lib.check(lib.svanon_set_gemm_mode(mode))
    try:
        _lib.check(lib.svanon_debug_gemm_taps(eng.handle, ptr(a_base), rows - 1, lda, margin, step, ptr(W), len(taps), offs, ptr(b),
                                              ptr(Cbuf), ldc, c0, M, N, K, None))
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.svanon_set_gemm_mode(2))
    flat = A.reshape(-1).double()
    ref = b.double().repeat(M, 1)
    for t, off in enumerate(taps):
        starts = (margin + torch.arange(M, device="cuda") * step + off) * lda + a_col0
        idx = starts[:, None] + torch.arange(K, device="cuda")[None]
        ref += flat[idx] @ W[t].double().T
    got = Cbuf[:, c0:c0 + N].double()
    assert float((got - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    untouched = torch.cat([Cbuf[:, :c0], Cbuf[:, c0 + N:]], dim=1)
    assert bool((untouched == 7.0).all())                                        # neighbours of the column slice survive

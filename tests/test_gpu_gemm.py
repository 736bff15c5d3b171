"""GEMM back ends against torch fp64 on the GPU: fp32 CUDA-core kernels and the tcgen05 3xTF32 kernel."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, M, N, K, act=0, seed=0):
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    eng = Engine.get(0)
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    _lib.check(lib.svanon_set_gemm_mode(mode))
    try:
        _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, act, None))
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.svanon_set_gemm_mode(2))
    ref = A.double() @ W.double().T + b.double()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    return out, ref


SHAPES = [(512, 1536, 384), (512, 384, 1536), (128, 1536, 512), (545, 2304, 768), (512, 2050, 2048), (100, 72, 48),
          (32, 256, 2816), (2048, 16, 48)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_simt_gemm(M, N, K):
    out, ref = _run(1, M, N, K)
    assert float((out.double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("M,N,K", SHAPES[:6])
def test_tcgen05_3xtf32_gemm(M, N, K):
    """fp32-grade accuracy: the 3xTF32 split must be as close to fp64 as the fp32 FMA kernel is (single-pass TF32
    would be ~1e-3)."""
    out, ref = _run(2, M, N, K)
    err = float((out.double() - ref).abs().max())
    assert err < 2e-5 * max(1.0, float(ref.abs().max())), err


@pytest.mark.parametrize("M,N,K", [(16384, 2048, 512), (16384, 512, 2048), (20992, 128, 512), (20000, 384, 1008),
                                   (16500, 300, 784)])
def test_tcgen05_wide_tiles(M, N, K):
    """Shapes large enough for the 128 x 256 and 128 x 128 CTA tiles (many-stream batches), ragged M / N / K."""
    out, ref = _run(2, M, N, K)
    err = float((out.double() - ref).abs().max())
    assert err < 2e-5 * max(1.0, float(ref.abs().max())), err


def test_tcgen05_gelu_epilogue():
    out, ref = _run(2, 512, 1536, 384, act=1)
    assert float((out.double() - ref).abs().max()) < 3e-5


@pytest.mark.parametrize("M,N,K", [(512, 1536, 384), (128, 1536, 512), (545, 2304, 768), (100, 72, 48), (16384, 2048, 512),
                                   (16384, 512, 2048), (20992, 128, 512), (16500, 300, 784)])
def test_tcgen05_fp16_perf_mode_gemm(M, N, K):
    """PERF mode (svanon_set_precision 1): one kind::f16 pass.  The kernel must compute exactly the product of the
    fp16-ROUNDED operands with fp32 accumulation -- checked against fp64 on the rounded operands (1e-5 relative: only the
    accumulation order differs) -- which sits ~1e-3 from the fp32 product, the price of the mode (printed)."""
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    eng, lib = Engine.get(0), _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    _lib.check(lib.svanon_set_precision(1))
    try:
        _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 0, None))
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.svanon_set_precision(0))
    rounded = A.half().double() @ W.half().double().T + b.double()
    exact = A.double() @ W.double().T + b.double()
    scale = max(1.0, float(exact.abs().max()))
    err_kernel = float((out.double() - rounded).abs().max())
    err_mode = float((out.double() - exact).abs().max())
    print(f"[perf mode] {M}x{N}x{K}: vs fp16-rounded operands {err_kernel:.2e}, vs fp32 product {err_mode:.2e} (scale {scale:.1f})")
    assert err_kernel < 1e-5 * scale, err_kernel
    assert err_mode < 2e-2 * scale, err_mode


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("M,N,K", [(512, 1536, 384), (128, 1536, 512), (545, 2304, 768), (100, 72, 48), (328, 520, 2048),
                                   (16384, 2048, 512), (16384, 512, 2048), (20992, 128, 512), (16500, 300, 784),
                                   (16384, 32, 224), (16384, 16, 112), (9216, 32, 96), (8192, 16, 48)])   # thin 128 x N tiles
def test_tcgen05_weights_by_tma(M, N, K, precision):
    """The engine's weight path: W pre-split (hi / lo) or converted (fp16) and pre-tiled in HBM, the B operand of every K-slab
    arriving by cp.async.bulk on the stage's mbarrier (ragged N / K: zero-filled tile rows).  Same bounds as the register path."""
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    eng, lib = Engine.get(0), _lib.load()
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    _lib.check(lib.svanon_set_precision(precision))
    _lib.check(lib.svanon_debug_gemm_weights_static(1))
    _lib.check(lib.svanon_set_gemm_pair(0))          # this test holds the single-CTA kernel; the pair kernel has its own below
    try:
        _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, 0, None))
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.svanon_set_gemm_pair(-1))
        _lib.check(lib.svanon_debug_gemm_weights_static(0))
        _lib.check(lib.svanon_set_precision(0))
    exact = A.double() @ W.double().T + b.double()
    scale = max(1.0, float(exact.abs().max()))
    if precision == 0:
        assert float((out.double() - exact).abs().max()) < 2e-5 * scale
    else:
        rounded = A.half().double() @ W.half().double().T + b.double()
        assert float((out.double() - rounded).abs().max()) < 1e-5 * scale


@pytest.mark.parametrize("M,N,K,act", [(16384, 2048, 512, 1), (16384, 512, 2048, 0), (20992, 384, 1536, 0), (20992, 128, 512, 0),
                                       (12500, 256, 96, 0), (16384, 1536, 512, 0)])
def test_pair_gemm_bitwise(M, N, K, act):
    """The CTA-pair kernel (csrc/gemm_pair.cu: tcgen05.mma.cta_group::2, both operands by tensor-map TMA, RAW fp32 arrays as the
    hi terms) computes the same bits as the single-CTA kernel (explicitly masked hi terms, operands through registers), with
    its own lo-split pass (mode 1) and with explicitly masked hi copies (mode 2); ragged M; the launch counter proves the
    kernel under test ran."""
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    eng, lib = Engine.get(0), _lib.load()
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    outs = {}
    _lib.check(lib.svanon_debug_gemm_weights_static(1))
    try:
        for mode in (0, 1, 2):
            out = torch.full((M, N), float("nan"), device="cuda")
            _lib.check(lib.svanon_set_gemm_pair(mode))
            n0 = lib.svanon_gemm_pair_launches()
            _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, act, None))
            torch.cuda.synchronize()
            assert lib.svanon_gemm_pair_launches() - n0 == (1 if mode else 0)
            outs[mode] = out
    finally:
        _lib.check(lib.svanon_set_gemm_pair(-1))
        _lib.check(lib.svanon_debug_gemm_weights_static(0))
    exact = A.double() @ W.double().T + b.double()
    if act == 1:
        exact = torch.nn.functional.gelu(exact)
    scale = max(1.0, float(exact.abs().max()))
    assert float((outs[1].double() - exact).abs().max()) < 2e-5 * scale
    assert torch.equal(outs[1], outs[0]), int((outs[1] != outs[0]).sum())
    assert torch.equal(outs[2], outs[0]), int((outs[2] != outs[0]).sum())


def _fused(lib, eng, ptr, A, W, W2, table, rope_cols, seg, M, N, K):
    from streamvoiceanon_b200 import _lib
    out = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.svanon_debug_gemm_fused(eng.handle, ptr(A), ptr(W), ptr(W2) if W2 is not None else None,
                                           ptr(table) if table is not None else None, rope_cols, seg, ptr(out), M, N, K, None))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("form", ["swiglu", "rope"])
def test_pair_gemm_fused_epilogues(form):
    """The fused forms of the many-stream encoder transformer: SwiGLU gate (w1 | w3 as the two halves of the pair's B tile, gate
    in the epilogue) and RoPE on q | k in the qkv epilogue -- the same bits as the GEMM followed by the row-wise kernels, and
    close to fp64."""
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    eng, lib = Engine.get(0), _lib.load()
    g = torch.Generator(device="cuda").manual_seed(23)
    M, K = 16384, 512
    A = torch.randn(M, K, device="cuda", generator=g)
    if form == "swiglu":
        N = 1536
        W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        W2 = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        args = (W, W2, None, 0, 0)
        h1, h3 = A.double() @ W.double().T, A.double() @ W2.double().T
        exact = torch.nn.functional.silu(h1) * h3
    else:
        N, S = 1536, 128
        W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        ang = torch.rand(S, 32, device="cuda", generator=g) * 6.28
        table = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()          # [pos][32][2]
        args = (W, None, table, 1024, S)
        y = (A.double() @ W.double().T).view(M, N)
        pos = torch.arange(M, device="cuda") % S
        c, s_ = table[pos, :, 0].double(), table[pos, :, 1].double()               # [M][32]
        qk = y[:, :1024].reshape(M, 16, 32, 2)
        x0, x1 = qk[..., 0], qk[..., 1]
        rot = torch.stack([x0 * c[:, None] - x1 * s_[:, None], x1 * c[:, None] + x0 * s_[:, None]], dim=-1).reshape(M, 1024)
        exact = torch.cat([rot, y[:, 1024:]], dim=1)
    outs = {}
    _lib.check(lib.svanon_debug_gemm_weights_static(1))
    try:
        for mode in (0, 1):
            _lib.check(lib.svanon_set_gemm_pair(mode))
            n0 = lib.svanon_gemm_pair_launches()
            outs[mode] = _fused(lib, eng, ptr, A, *args, M, N, K)
            assert lib.svanon_gemm_pair_launches() - n0 == mode
    finally:
        _lib.check(lib.svanon_set_gemm_pair(-1))
        _lib.check(lib.svanon_debug_gemm_weights_static(0))
    scale = max(1.0, float(exact.abs().max()))
    assert float((outs[1].double() - exact).abs().max()) < 3e-5 * scale
    assert torch.equal(outs[1], outs[0]), int((outs[1] != outs[0]).sum())


@pytest.mark.parametrize("M,N,C,offs,k1", [(12544, 128, 128, (-10, -5, 0), 0), (12500, 256, 256, (-30, -25, -20, -15, -10, -5, 0), 0),
                                           (12544, 256, 256, (-2,), 3), (12544, 128, 64, (-6,), 7), (16384, 64, 64, (-20, -10, 0), 0), (16384, 32, 32, (-10, -5, 0), 0),
                                           (16384, 32, 32, (-6,), 7), (16384, 16, 16, (-10, -5, 0), 0), (16384, 16, 16, (-10,), 11)])
def test_pair_gemm_taps_bitwise(M, N, C, offs, k1):
    """The conv form of the pair kernel (causal convs as GEMMs over taps: 3-D tensor maps over the input rows with their left
    context, tap t at row offset tap_off[t]; k1 > 0: the dilation-1 form, one "tap" of k1 overlapping rows) against the single-CTA
    kernel: the same bits, and fp64-close."""
    import ctypes as C_
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    eng, lib = Engine.get(0), _lib.load()
    g = torch.Generator(device="cuda").manual_seed(31)
    taps = len(offs)
    K = k1 * C if k1 else C
    margin = -min(offs)
    rows = margin + M + (k1 if k1 else 0) + 1
    A = torch.randn(rows, C, device="cuda", generator=g)
    W = torch.randn(taps, N, K, device="cuda", generator=g) / (taps * K) ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    arr = (C_.c_int * taps)(*offs)
    outs = {}
    _lib.check(lib.svanon_debug_gemm_weights_static(1))
    try:
        for mode in (0, 1):
            out = torch.full((M, N), float("nan"), device="cuda")
            _lib.check(lib.svanon_set_gemm_pair(mode))
            n0 = lib.svanon_gemm_pair_launches()
            _lib.check(lib.svanon_debug_gemm_taps(eng.handle, ptr(A), rows - 1, C, margin, 1, ptr(W), taps, arr, ptr(b), ptr(out), N, 0,
                                                  M, N, K, None))
            torch.cuda.synchronize()
            assert lib.svanon_gemm_pair_launches() - n0 == mode
            outs[mode] = out
    finally:
        _lib.check(lib.svanon_set_gemm_pair(-1))
        _lib.check(lib.svanon_debug_gemm_weights_static(0))
    flat = A.reshape(-1).double()
    ref = b.double().repeat(M, 1)
    for t, off in enumerate(offs):
        starts = (margin + torch.arange(M, device="cuda") + off) * C
        idx = starts[:, None] + torch.arange(K, device="cuda")[None]
        ref += flat[idx] @ W[t].double().T
    assert float((outs[1].double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    assert torch.equal(outs[1], outs[0]), int((outs[1] != outs[0]).sum())

"""The persistent chain kernel (csrc/chain.cu: op-list interpreter, tcgen05 GEMM phases + row-wise phases, grid barriers
instead of launches) against (1) torch fp64 for its GEMM phase, (2) the per-op path for the encoder transformer, on the
same inputs, and (3) the per-op path over a streaming session (ids bit-exact, wave within 1e-10).  The reference fixtures
(tests/test_gpu_parity.py) run through the chain as well, because it is the default single-stream path."""
import os

import pytest
import torch

from streamvoiceanon_b200 import synth

pytestmark = pytest.mark.gpu


def _lib_eng():
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    return _lib, _lib.load(), Engine.get(0), ptr


# the shapes the single-stream encoder issues (transformer at 128 tokens and at the kept tail, ConvNeXt stages on 2 x 164
# rows, the down-sampling blocks on 164 / 82 rows) plus ragged ones
CHAIN_SHAPES = [(128, 1536, 512), (128, 512, 512), (128, 512, 1536), (1, 512, 512), (8, 1536, 512),
                (2, 512, 1536), (328, 512, 128), (328, 128, 512), (328, 1024, 256), (328, 256, 1024), (328, 1536, 384),
                (328, 384, 1536), (328, 2048, 512), (328, 512, 2048), (164, 512, 1024), (82, 2048, 512), (82, 512, 2048),
                (100, 48, 64), (384, 80, 96), (257, 16, 32)]


@pytest.mark.parametrize("repeat", [1, 3])
@pytest.mark.parametrize("M,N,K", CHAIN_SHAPES)
def test_chain_gemm_phase_vs_fp64(M, N, K, repeat):
    """One GEMM phase (weights by TMA into shared memory a GEMM ahead, A through the producer warps, 3xTF32 tcgen05.mma into
    one TMEM accumulator per M tile, K-slice partials to global memory) + the element-wise phase that sums the partials:
    the same 2e-5 bound against fp64 as the stand-alone GEMM kernels.  repeat = 3 runs three such pairs in one launch
    (weight double buffer, barrier parities, TMEM reuse)."""
    _lib, lib, eng, ptr = _lib_eng()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    for act in (0, 1):
        out = torch.full((M, N), float("nan"), device="cuda")
        _lib.check(lib.svanon_debug_chain_gemm(eng.handle, ptr(A), ptr(W), ptr(b), ptr(out), M, N, K, act, repeat, None))
        torch.cuda.synchronize()
        ref = A.double() @ W.double().T + b.double()
        if act:
            ref = torch.nn.functional.gelu(ref)
        err = float((out.double() - ref).abs().max())
        assert err < 2e-5 * max(1.0, float(ref.abs().max())), (err, act)


@pytest.mark.parametrize("S,keep", [(128, 1), (128, 2), (128, 0), (96, 4), (40, 1), (128, 8)])
def test_chain_encoder_transformer_equals_per_op_path(models, S, keep):
    """WindowLimitedTransformer + BSQ of one window (windowed_transformer.py:337-354, bsq.py:330-369) as one chain launch
    against the per-op kernels (which the reference fixtures pin): final-norm rows within 1e-4 after 8 layers (values are
    O(1); the two paths sum the K slices in different orders and the chain applies RoPE / SiLU-mul in the GEMM epilogue; a
    wrong attention or MLP output moves the rows by >= 1e-2 through the layer scales), ids identical."""
    _lib, lib, eng, ptr = _lib_eng()
    g = torch.Generator(device="cuda").manual_seed(S * 31 + keep)
    xt = torch.randn(S, 512, device="cuda", generator=g)
    rows = keep if keep > 0 else S
    res = []
    for use_chain in (0, 1):
        hid = torch.full((rows, 512), float("nan"), device="cuda")
        ids = torch.full((S,), -1, dtype=torch.int64, device="cuda")
        _lib.check(lib.svanon_debug_enc_transformer(eng.handle, ptr(xt), S, keep, use_chain, ptr(hid), ptr(ids), None))
        torch.cuda.synchronize()
        res.append((hid.cpu(), ids.cpu()))
    assert torch.isfinite(res[1][0]).all()
    err = float((res[0][0] - res[1][0]).abs().max())
    assert err < 1e-4, err
    assert torch.equal(res[0][1], res[1][1])
    assert (res[1][1][S - rows:] >= 0).all()


def test_chain_stream_loop_equals_per_op_path(models, gold, tape):
    """The streaming loop with the encoder's transformer half as a chain launch (default), and with the conv stack inside
    the chain as well (mode 3), against the same loop with one kernel launch per op: content ids, codec ids identical,
    waveform within 1e-10, over 40 chunks (window state carried by the chain's assemble phase)."""
    from streamvoiceanon_b200 import StreamSession
    _lib, lib, eng, ptr = _lib_eng()
    g = gold("stream_default")
    n_ref = int(g["n_ref"])
    style, timbre = synth.synth_speaker(int(g["ref_seed"]))
    gen = torch.Generator().manual_seed(int(g["codes_seed"]))
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    ref_content = torch.from_numpy(g["ref_content"])
    n_chunks = 40
    src = synth.synth_audio_44k(1301, 3.0)[: n_chunks * 2048].view(n_chunks, 2048)
    out = []
    try:
        for chain in (1, 3, 0):                           # transformer half; + conv stack; one launch per op
            _lib.check(lib.svanon_set_chain_mode(chain))
            sess = StreamSession()
            sess.set_noise_fn(tape(7601), 0)
            sess.set_prompt(ref_content[0].cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda(), 256, 2)
            sess.setup(128, 64, 768, 32, 1)
            waves = torch.cat([sess.process_chunk(src[i].cuda()).cpu() for i in range(n_chunks)])
            out.append((*sess.history(), waves))
            sess.close()
    finally:
        _lib.check(lib.svanon_set_chain_mode(int(os.environ.get("SVANON_CHAIN", "1"))))
    for k in (0, 1):
        assert torch.equal(out[k][0], out[2][0]), k
        assert torch.equal(out[k][1], out[2][1]), k
        assert float(((out[k][2] - out[2][2]) ** 2).mean()) < 1e-10, k

"""GPU parity tests: the CUDA path (through the C ABI, via the reference-facing shims) against
 (1) the golden fixtures produced by the UNMODIFIED reference (tests/golden, oracle/make_golden.py), and
 (2) the CPU oracle on the same seeded inputs.
Bar: bit-exact for every integer output (BSQ content ids, codec ids, positions); floating point within the
tolerance written beside each check (fp32 on both sides; only summation order / libm differ)."""
import numpy as np
import pytest
import torch

from streamvoiceanon_b200 import synth

pytestmark = pytest.mark.gpu

WAVE_MSE_TOL = 1e-8          # fp32 waveform MSE tolerance
LOGIT_TOL = 2e-3             # max-abs error of teacher-forced logits (values are O(3))


def test_native_library_is_loaded(models):
    import os
    maps = open(f"/proc/{os.getpid()}/maps").read()
    assert "libsvanon_b200.so" in maps
    from streamvoiceanon_b200 import _lib
    assert _lib.kernel_launches() >= 0


# ------------------------------------------------------------------------------------------------ E
def test_encoder_40_frames_vs_reference(models, gold):
    _, tok, _ = models
    g = gold("encoder_40f")
    wav = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: int(g["n_samples"])][None]
    ids, flen = tok.encode(wav.cuda(), torch.LongTensor([wav.shape[1]]).cuda())
    assert int(flen[0]) == 40
    assert np.array_equal(ids.cpu().numpy(), g["ids"])


def test_encoder_host_buffers(models, gold):
    """Same call with HOST tensors: the C ABI stages them itself."""
    _, tok, _ = models
    g = gold("encoder_40f")
    wav = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: int(g["n_samples"])][None]
    ids, _ = tok.encode(wav, torch.LongTensor([wav.shape[1]]))
    assert np.array_equal(ids.cpu().numpy(), g["ids"])


def test_encoder_streaming_window_vs_reference(models, gold):
    _, tok, _ = models
    g = gold("encoder_window128")
    live = int(g["live_frames"])
    win = torch.zeros(1, 128 * 2048)
    win[:, -live * 2048:] = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: live * 2048]
    ids, _ = tok.encode(win.cuda(), torch.LongTensor([win.shape[1]]).cuda())
    assert np.array_equal(ids.cpu().numpy(), g["ids"])


def test_encoder_vs_oracle_other_inputs(models, weights):
    """Seeds and lengths the fixtures do not cover, incl. ragged lengths and the one-frame minimum."""
    from oracle import content_encoder as E
    _, tok, _ = models
    for seed, n in ((2001, 2048), (2002, 13 * 2048 + 700), (2003, 61 * 2048)):
        wav = synth.synth_audio_44k(seed, 3.0)[:n][None]
        with torch.no_grad():
            ref, _ = E.encode(wav, weights["tok"])
        ids, _ = tok.encode(wav.cuda(), torch.LongTensor([n]).cuda())
        assert np.array_equal(ids.cpu().numpy(), ref.numpy()), (seed, n)


def test_encoder_prefix_causality(models):
    """Property (size-independent): ids of x[:N] are the first N ids of x."""
    _, tok, _ = models
    wav = synth.synth_audio_44k(2004, 3.0)[: 60 * 2048][None].cuda()
    full, _ = tok.encode(wav, torch.LongTensor([wav.shape[1]]).cuda())
    part, _ = tok.encode(wav[:, : 33 * 2048].contiguous(), torch.LongTensor([33 * 2048]).cuda())
    assert torch.equal(full[0, 0, :33], part[0, 0])


def test_encoder_rejects_too_short(models):
    _, tok, _ = models
    ids, _ = tok.encode(torch.zeros(1, 1000).cuda(), torch.LongTensor([1000]).cuda())
    assert ids.shape[-1] == 0


# ------------------------------------------------------------------------------------------------ V
def test_vocoder_vs_reference(models, gold):
    _, _, voc = models
    g = gold("vocoder_20f")
    codes = torch.from_numpy(g["codes"]).cuda()
    z = voc.quantizer.decode(codes)
    assert z.shape == (1, 512, 80)
    assert np.abs(z[0, :, -8:].cpu().numpy() - g["z_tail"]).max() < 1e-4
    wave = voc.head(z)
    assert wave.shape == (1, 1, 20 * 2048)
    mse = float(((wave[0, 0].cpu().numpy() - g["wave"]) ** 2).mean())
    assert mse < WAVE_MSE_TOL, mse
    fused = voc.decode_codes(codes)
    assert torch.equal(fused, wave)


def test_vocoder_vs_oracle_and_causality(models, weights):
    from oracle import vocoder as V
    _, _, voc = models
    g = torch.Generator().manual_seed(77)
    codes = torch.randint(0, 1000, (1, 8, 9), generator=g)
    with torch.no_grad():
        ref = V.code2wav(codes, weights["voc_folded"])
    wave = voc.decode_codes(codes.cuda())
    assert float(((wave.cpu() - ref) ** 2).mean()) < WAVE_MSE_TOL
    # strict causality: changing the last frame leaves every earlier sample bit-identical
    codes2 = codes.clone()
    codes2[:, :, -1] = (codes2[:, :, -1] + 17) % 1000
    wave2 = voc.decode_codes(codes2.cuda())
    assert torch.equal(wave[..., : 8 * 2048], wave2[..., : 8 * 2048])
    assert not torch.equal(wave[..., 8 * 2048:], wave2[..., 8 * 2048:])


def test_vocoder_window_receptive_field(models):
    """the last frame depends on the last 16 code frames only, so a 16-frame window
    reproduces the tail of a 64-frame window (the property the stateful design relies on)."""
    _, _, voc = models
    g = torch.Generator().manual_seed(78)
    codes = torch.randint(0, 1000, (1, 8, 64), generator=g).cuda()
    w64 = voc.decode_codes(codes)[..., -2048:]
    w16 = voc.decode_codes(codes[:, :, -16:].contiguous())[..., -2048:]
    assert float(((w64 - w16) ** 2).mean()) < 1e-10


# ------------------------------------------------------------------------------------------------ A
def _prefill(ar, s, tape):
    style, timbre = synth.synth_speaker(int(s["spk_seed"]))
    ar.set_delay(delay=int(s["delay"]))
    ar.set_noise_fn(tape(int(s["tape_seed"])), 0)
    ar.prefill_prompt(torch.from_numpy(s["ref_content"]).cuda(), torch.from_numpy(s["ref_audio"]).cuda(),
                      style.cuda(), timbre.cuda())
    return torch.from_numpy(s["src_content"])


def test_ar_streaming_codes_vs_reference(models, gold, tape):
    ar, _, _ = models
    s = gold("ar_stream")
    src = _prefill(ar, s, tape)
    ar.prefill_src_condition4delay(src[:, :2].cuda())
    for i, t in enumerate(range(2, 16)):
        codes, pos = ar.decode_one(src[:, t:t + 1].cuda())
        assert codes.dtype == torch.int32 and tuple(codes.shape) == (8, 1)
        assert np.array_equal(codes.cpu().numpy(), s["codes"][i]), (i, codes.T.tolist(), s["codes"][i].T.tolist())
        assert int(pos) == int(s["pos"][i])


@pytest.mark.parametrize("variant", [0])
def test_ar_other_kernel_variants(models, gold, tape, variant):
    """The other batch-1 kernel (0: weights straight from global memory, the kernel 2- and 4-stream launches use)
    produces the same codes as the default (1: TMA-staged weights)."""
    from streamvoiceanon_b200 import _lib
    ar, _, _ = models
    s = gold("ar_stream")
    _lib.check(_lib.load().svanon_ar_set_kernel_variant(ar._engine.handle, variant))
    try:
        src = _prefill(ar, s, tape)
        ar.prefill_src_condition4delay(src[:, :2].cuda())
        for i, t in enumerate(range(2, 8)):
            codes, pos = ar.decode_one(src[:, t:t + 1].cuda())
            assert np.array_equal(codes.cpu().numpy(), s["codes"][i]), i
    finally:
        _lib.check(_lib.load().svanon_ar_set_kernel_variant(ar._engine.handle, 1))


def test_ar_teacher_forced_logits_vs_reference(models, gold, tape):
    ar, _, _ = models
    s, g = gold("ar_stream"), gold("ar_logits")
    src = _prefill(ar, s, tape)
    ar.prefill_src_condition4delay(src[:, :2].cuda())
    ar.debug_logits(True)
    try:
        codes, _ = ar.decode_one(src[:, 2:3].cuda())
        slow, hidden, fast = ar.read_debug()
    finally:
        ar.debug_logits(False)
    assert np.abs(hidden.numpy() - g["hidden"]).max() < LOGIT_TOL
    assert np.abs(slow.numpy() - g["slow_logits"]).max() < LOGIT_TOL
    assert np.abs(fast.numpy() - g["fast_logits"]).max() < LOGIT_TOL
    assert np.array_equal(codes.cpu().numpy(), g["codes"])


def test_ar_offline_generate_vs_reference(models, gold, tape):
    ar, _, _ = models
    s, g = gold("ar_stream"), gold("ar_generate")
    style, timbre = synth.synth_speaker(int(s["spk_seed"]))
    ar.set_delay(delay=2)
    ar.set_noise_fn(tape(int(s["tape_seed"])), 0)
    out = ar.generate(torch.from_numpy(s["ref_content"]).cuda(), torch.from_numpy(s["ref_audio"]).cuda(),
                      torch.from_numpy(s["src_content"])[:, :10].cuda(), style.cuda(), timbre.cuda())
    assert np.array_equal(out.cpu().numpy(), g["codes"])


def test_ar_vs_oracle_long_prompt_other_delay(models, weights, tape):
    """A longer prompt (S_valid ~ 400), delay 4, different tape: 10 frames bit-exact vs the oracle."""
    from oracle.dual_ar import DualAR
    ar, _, _ = models
    g = torch.Generator().manual_seed(555)
    T = 180
    ref_content = torch.randint(0, 8192, (1, T), generator=g)
    ref_audio = torch.randint(0, 1000, (1, 8, T), generator=g).int()
    src = torch.randint(0, 8192, (1, 14), generator=g)
    style, timbre = synth.synth_speaker(5003)
    orc = DualAR(weights["ar"], tape(7100))
    with torch.no_grad():
        orc.set_delay(4)
        orc.prefill_prompt(ref_content, ref_audio, style, timbre)
        orc.prefill_src_condition4delay(src[:, :4])
        want = [orc.decode_one(src[:, t:t + 1]) for t in range(4, 14)]
    ar.set_delay(delay=4)
    ar.set_noise_fn(tape(7100), 0)
    ar.prefill_prompt(ref_content.cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda())
    ar.prefill_src_condition4delay(src[:, :4].cuda())
    for i, t in enumerate(range(4, 14)):
        codes, pos = ar.decode_one(src[:, t:t + 1].cuda())
        assert np.array_equal(codes.cpu().numpy(), want[i][0].numpy()), i
        assert int(pos) == int(want[i][1])


def test_ar_batched_streams_match_single_stream(models, weights, tape):
    """Engine-only property (the reference is batch-1): stream i of a 2-stream launch == stream i alone."""
    import ctypes as C
    from streamvoiceanon_b200 import ARVCWrapper, _lib
    from streamvoiceanon_b200.engine import ptr
    ar0, _, _ = models
    g = torch.Generator().manual_seed(900)
    prompts = []
    for b, T in enumerate((20, 37)):
        prompts.append((torch.randint(0, 8192, (1, T), generator=g), torch.randint(0, 1000, (1, 8, T), generator=g).int(),
                        torch.randint(0, 8192, (1, 8), generator=g), synth.synth_speaker(6000 + b)))
    singles = []
    for b, (rc, ra, src, (style, timbre)) in enumerate(prompts):
        ar0.set_delay(delay=2)
        ar0.set_noise_fn(tape(7200 + b), 0)
        ar0.prefill_prompt(rc.cuda(), ra.cuda(), style.cuda(), timbre.cuda())
        ar0.prefill_src_condition4delay(src[:, :2].cuda())
        singles.append([ar0.decode_one(src[:, t:t + 1].cuda())[0].cpu() for t in range(2, 8)])
    wrappers = []
    for b, (rc, ra, src, (style, timbre)) in enumerate(prompts):
        w = ARVCWrapper()
        w.setup_caches(max_batch_size=1, max_seq_len=2048)
        w.set_delay(delay=2)
        w.prefill_prompt(rc.cuda(), ra.cuda(), style.cuda(), timbre.cuda())
        w.prefill_src_condition4delay(src[:, :2].cuda())
        wrappers.append(w)
    lib = _lib.load()
    handles = (C.c_void_p * 2)(wrappers[0]._stream, wrappers[1]._stream)
    for i, t in enumerate(range(2, 8)):
        ids = torch.tensor([int(prompts[0][2][0, t]), int(prompts[1][2][0, t])], dtype=torch.int64).cuda()
        noise = torch.stack([torch.stack([tape(7200 + b)(2 + i, s, 1000)[:1000] for s in range(1, 9)]) for b in range(2)])
        noise = noise.float().contiguous().cuda()
        out = torch.empty(2, 8, dtype=torch.int32, device="cuda")
        _lib.check(lib.svanon_ar_decode_batch(handles, 2, ptr(ids), ptr(noise), ptr(out), None))
        for b in range(2):
            assert torch.equal(out[b].cpu(), singles[b][i][:, 0]), (b, i)


def test_ar_errors(models):
    ar, _, _ = models
    with pytest.raises(AssertionError):
        ar.set_delay(delay=2)
        ar.prefill_src_condition4delay(torch.zeros(1, 5, dtype=torch.long).cuda())     # != delay (dual_ar_stream.py:805)
    with pytest.raises(ValueError):
        ar.setup_caches(max_batch_size=2, max_seq_len=2048)


# ------------------------------------------------------------------------------------------------ loop
def _run_loop(models, weights, g, tape, host_io, incremental=True):
    from streamvoiceanon_b200 import StreamSession
    _, tok, _ = models
    n_ref, n_chunks = int(g["n_ref"]), int(g["n_chunks"])
    style, timbre = synth.synth_speaker(int(g["ref_seed"]))
    ref_wave = synth.synth_audio_44k(int(g["ref_seed"]), 3.5)[: n_ref * 2048][None]
    gen = torch.Generator().manual_seed(int(g["codes_seed"]))
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    assert np.array_equal(ref_content[0].cpu().numpy(), g["ref_content"])
    sess = StreamSession()
    sess.set_noise_fn(tape(int(g["tape_seed"])), 0)
    sess.set_prompt(ref_content[0].cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda(), max_prompt_frames=256,
                    delay=int(g["delay"]))
    sess.setup(int(g["encode_window_frames"]), int(g["decode_window_frames"]), int(g["max_seq_frames"]),
               int(g["buffer_frames"]), 1)
    sess.set_vocoder_mode(incremental)
    src = synth.synth_audio_44k(int(g["src_seed"]), 1.5)[: n_chunks * 2048].view(n_chunks, 2048)
    waves = []
    for i in range(n_chunks):
        chunk = src[i] if host_io else src[i].cuda()
        waves.append(sess.process_chunk(chunk).cpu())
    src_hist, pred_hist = sess.history()
    sess.close()
    return src_hist, pred_hist, torch.cat(waves)


def test_incremental_vocoder_equals_window_recompute(models, weights, gold, tape):
    """The incremental vocoder (new frames only, per-stream conv history) against the reference-style recompute
    of the 64-frame window inside the same loop: identical codec ids, waveform equal to fp32 rounding."""
    g = gold("stream_default")
    _, pred_a, wave_a = _run_loop(models, weights, g, tape, False, incremental=True)
    _, pred_b, wave_b = _run_loop(models, weights, g, tape, False, incremental=False)
    assert torch.equal(pred_a, pred_b)
    assert float(((wave_a - wave_b) ** 2).mean()) < 1e-10
    assert float((wave_a - wave_b).abs().max()) < 1e-4


@pytest.mark.parametrize("name,host_io,incremental", [("stream_reprompt", False, True), ("stream_default", True, True),
                                                      ("stream_default", False, False)])
def test_stream_loop_vs_reference(models, weights, gold, tape, name, host_io, incremental):
    """The whole per-chunk loop against the UNMODIFIED reference's process_one_chunk: content ids and codec ids
    bit-exact, waveform within the fp32 MSE tolerance.  `stream_default` = CLI defaults (windows 128/64) with
    HOST buffers through the C ABI; `stream_reprompt` = small windows with the re-prompt path firing."""
    g = gold(name)
    src_hist, pred_hist, wave = _run_loop(models, weights, g, tape, host_io, incremental)
    assert np.array_equal(src_hist.numpy()[None], g["src_content"])
    assert np.array_equal(pred_hist.numpy()[None], g["pred_codes"])
    mse = float(((wave.numpy() - g["wave"]) ** 2).mean())
    assert mse < WAVE_MSE_TOL, mse


def test_stream_loop_chunk2_truncated_prompt_vs_oracle(models, weights, tape):
    """BASELINE config 5 shape: decode_chunk_frames=2, delay=2, a prompt longer than max_prompt_frames (the
    reference prefills the UNtruncated prompt but pads / re-prompts with the truncated copy, infer_arvc.py:469-489),
    re-prompt firing.  Content ids and codec ids bit-exact vs the CPU oracle, waveform within the fp32 tolerance."""
    from oracle import content_encoder as E
    from oracle.streaming import StreamOracle
    from streamvoiceanon_b200 import StreamSession
    _, tok, _ = models
    n_ref, n_chunks, chunk, max_prompt = 40, 9, 2, 30
    cfg = dict(encode_window_frames=24, decode_window_frames=24, max_seq_frames=66, buffer_frames=6,
               decode_chunk_frames=chunk)
    style, timbre = synth.synth_speaker(5200)
    ref_wave = synth.synth_audio_44k(5200, 3.0)[: n_ref * 2048][None]
    gen = torch.Generator().manual_seed(321)
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    src = synth.synth_audio_44k(1200, 1.5)[: n_chunks * chunk * 2048].view(n_chunks, chunk * 2048)
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    sess = StreamSession()
    sess.set_noise_fn(tape(7500), 0)
    sess.set_prompt(ref_content[0].cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda(), max_prompt_frames=max_prompt, delay=2)
    sess.setup(**cfg)
    waves = torch.cat([sess.process_chunk(src[i].cuda()).cpu() for i in range(n_chunks)])
    src_hist, pred_hist = sess.history()
    sess.close()
    so = StreamOracle(weights["ar"], weights["tok"], weights["voc_folded"], tape(7500))
    with torch.no_grad():
        rc = E.encode(ref_wave, weights["tok"])[0].squeeze(0)
        so.prefill_prompt(ref_audio, rc, style, timbre, max_prompt, 2)
        so.setup_stream_caches(**cfg)
        want = torch.cat([so.process_one_chunk(src[i][None]) for i in range(n_chunks)], dim=-1)[0]
    assert np.array_equal(src_hist.numpy(), so.src_content_codes[0].numpy())
    assert np.array_equal(pred_hist.numpy(), so.pred_codes[0].numpy())
    assert float(((waves - want) ** 2).mean()) < WAVE_MSE_TOL


def test_ring_buffer_encoder_equals_window_recompute(models, weights, gold, tape):
    """E with persistent ring-buffer state (conv-stack outputs kept between chunks, only the window's first and last
    40 + chunk frames re-encoded) against the reference-style full re-encode of the 128-frame window in the same loop:
    identical content ids (and therefore codec ids) over 150 chunks (every state row has been produced incrementally by then); both modes are also held to the reference
    fixture by test_stream_loop_vs_reference."""
    from streamvoiceanon_b200 import StreamSession
    _, tok, _ = models
    g = gold("stream_default")
    n_ref = int(g["n_ref"])
    style, timbre = synth.synth_speaker(int(g["ref_seed"]))
    gen = torch.Generator().manual_seed(int(g["codes_seed"]))
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    ref_content = torch.from_numpy(g["ref_content"])
    n_chunks = 150                                  # > one full window turnover (128 frames) of the ring-buffer state
    src = synth.synth_audio_44k(1300, 8.0)[: n_chunks * 2048].view(n_chunks, 2048)
    out = []
    for incremental in (1, 0, 2):                  # ring-buffer state | full re-encode | + per-layer conv history
        sess = StreamSession()
        sess.set_noise_fn(tape(7600), 0)
        sess.set_prompt(ref_content[0].cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda(), 256, 2)
        sess.setup(128, 64, 768, 32, 1)
        sess.set_encoder_mode(incremental)
        waves = torch.cat([sess.process_chunk(src[i].cuda()).cpu() for i in range(n_chunks)])
        out.append((*sess.history(), waves))
        sess.close()
    for other in (out[1], out[2]):
        assert torch.equal(out[0][0], other[0])
        assert torch.equal(out[0][1], other[1])
        assert float(((out[0][2] - other[2]) ** 2).mean()) < 1e-10


# ------------------------------------------------------------------------------------------------ prompt path
def test_vocoder_encode_vs_reference(models, gold):
    """reference wave -> codec ids (`wav2target_fn`, infer_arvc.py:168-171) through
    svanon_voc_encode against the reference's own `FireflyArchitecture.encode`: ids bit-exact."""
    _, _, voc = models
    g = gold("vocoder_encode")
    for n in "ab":
        frames = int(g[f"frames_{n}"])
        wav = synth.synth_audio_44k(int(g[f"seed_{n}"]), 3.0)[: frames * 2048][None]
        (codes, _), lens = voc.encode(wav.cuda(), torch.LongTensor([wav.shape[1]]).cuda())
        assert codes.dtype == torch.int32 and tuple(codes.shape) == (1, 8, frames) and int(lens[0]) == frames
        assert np.array_equal(codes.cpu().numpy(), g[f"codes_{n}"]), n


def test_vocoder_encode_batched_and_ragged_vs_oracle(models, weights):
    """Rows side by side in one call, host buffers, and a ragged row (valid prefix only) against the oracle."""
    from oracle import vocoder as V
    _, _, voc = models
    n = 19 * 2048
    wavs = torch.stack([synth.synth_audio_44k(1500 + i, 1.0)[:n] for i in range(3)])
    with torch.no_grad():
        want = V.wav2codes(wavs, weights["voc_enc"])
    (codes, _), _ = voc.encode(wavs, torch.LongTensor([n] * 3))                     # host tensors
    assert np.array_equal(codes.cpu().numpy(), want.numpy())
    (ragged, _), lens = voc.encode(wavs.cuda(), torch.LongTensor([n, 7 * 2048 + 100, n]).cuda())
    assert lens.tolist() == [19, 7, 19]
    assert np.array_equal(ragged[1, :, :7].cpu().numpy(), want[1, :, :7].numpy())
    assert int(ragged[1, :, 7:].abs().sum()) == 0
    assert np.array_equal(ragged[0].cpu().numpy(), want[0].numpy())


def test_noise_mixing_vs_reference(models, gold):
    """svanon_noise_mix against the unmodified reference method
    (`InferenceWrapper.apply_noise_mixing`, infer_arvc.py:228-232) on the recorded draws, host and device buffers, and
    against the oracle; fp32, tolerance 1e-6 absolute.  With `noise=None` the shim draws `torch.randn_like` from the
    global generator where the reference does."""
    from oracle import prompt as P
    from streamvoiceanon_b200.prompt import apply_noise_mixing
    g = gold("noise_mix")
    for n in g["names"]:
        x, nz, alpha = torch.from_numpy(g[f"x_{n}"]), torch.from_numpy(g[f"noise_{n}"]), float(g[f"alpha_{n}"])
        y_host = apply_noise_mixing(x, alpha, nz)                                # host buffers through the C ABI
        y_dev = apply_noise_mixing(x.cuda(), alpha, nz.cuda())
        assert not y_host.is_cuda and y_dev.is_cuda and y_host.shape == x.shape
        assert np.abs(y_host.numpy() - g[f"y_{n}"]).max() < 1e-6, n
        assert np.array_equal(y_dev.cpu().numpy(), y_host.numpy()), n
        assert np.abs(y_host.numpy() - P.apply_noise_mixing(g[f"x_{n}"], alpha, g[f"noise_{n}"])).max() < 1e-6, n
    x = torch.from_numpy(g["x_style"]).cuda()
    torch.manual_seed(5)
    a = apply_noise_mixing(x, 0.7)
    torch.manual_seed(5)
    b = apply_noise_mixing(x, 0.7, torch.randn_like(x))
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        apply_noise_mixing(x, 0.7, torch.zeros(3))


# ------------------------------------------------------------------------------------------------ limits
def test_limits_kv_cache_full_and_maximum_lengths(models):
    """Maximum sizes: a full KV cache refuses to decode further (the reference would index out of range,
    dual_ar_stream.py:141-150), the tokenizer refuses more than the 2048 positions of its RoPE table, and the one-frame
    minimum works for both encoders and the vocoder."""
    from streamvoiceanon_b200 import ARVCWrapper
    _, tok, voc = models
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=64)
    ar.set_delay(delay=2)
    g = torch.Generator().manual_seed(3)
    T = 10                                                    # 33 + 2*10 = 53 prompt tokens, +3 for the delay prefill
    style, timbre = synth.synth_speaker(5300)
    ar.prefill_prompt(torch.randint(0, 8192, (1, T), generator=g).cuda(), torch.randint(0, 1000, (1, 8, T), generator=g).int().cuda(),
                      style.cuda(), timbre.cuda())
    ar.prefill_src_condition4delay(torch.randint(0, 8192, (1, 2), generator=g).cuda())
    for _ in range(4):                                        # positions 56..63
        ar.decode_one(torch.randint(0, 8192, (1, 1), generator=g).cuda())
    with pytest.raises(RuntimeError, match="KV cache full"):
        ar.decode_one(torch.randint(0, 8192, (1, 1), generator=g).cuda())
    with pytest.raises(RuntimeError, match="does not fit the KV cache"):
        ar.prefill_prompt(torch.zeros(1, 40, dtype=torch.long).cuda(), torch.zeros(1, 8, 40, dtype=torch.int32).cuda(),
                          style.cuda(), timbre.cuda())
    with pytest.raises(RuntimeError, match="RoPE table"):
        tok.encode(torch.zeros(1, 2049 * 2048).cuda(), torch.LongTensor([2049 * 2048]).cuda())
    one = synth.synth_audio_44k(1700, 0.2)[:2048][None].cuda()
    ids, flen = tok.encode(one, torch.LongTensor([2048]).cuda())
    assert tuple(ids.shape) == (1, 1, 1) and int(flen[0]) == 1
    (codes, _), _ = voc.encode(one, torch.LongTensor([2048]).cuda())
    assert tuple(codes.shape) == (1, 8, 1)
    wave = voc.decode_codes(codes.long())
    assert tuple(wave.shape) == (1, 1, 2048) and bool(torch.isfinite(wave).all())

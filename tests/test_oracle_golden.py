"""Pins the CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only.  Integer outputs must match exactly; floating-point
outputs to the tolerance written beside each check (both sides are fp32 on CPU, the
only differences are operator fusion / summation order)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import content_encoder as E
from oracle import vocoder as V
from oracle.dual_ar import DualAR
from oracle.streaming import StreamOracle
from streamvoiceanon_b200 import synth


def _digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes()[:4096])
    return h.hexdigest()


def test_synthetic_weights_are_reproducible(weights, gold):
    d = gold("weights_digest")
    assert _digest(weights["ar"]) == str(d["ar"])
    assert _digest(weights["tok"]) == str(d["tok"])
    assert _digest(weights["voc"]) == str(d["voc"])


def test_mel_filterbank_matches_torchaudio():
    import torchaudio.functional as AF
    fb = AF.melscale_fbanks(n_freqs=1025, f_min=0.0, f_max=22050.0, n_mels=160, sample_rate=44100,
                            norm="slaney", mel_scale="slaney")
    assert torch.allclose(E.slaney_fbanks(), fb, atol=1e-9, rtol=1e-6)


def test_encoder_40_frames(weights, gold):
    g = gold("encoder_40f")
    wav = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: int(g["n_samples"])][None]
    with torch.no_grad():
        ids, flen = E.encode(wav, weights["tok"])
        mel = E.log_mel(wav)
    assert int(flen[0]) == 40
    assert np.abs(mel[0, :, -8:].numpy() - g["mel_tail"]).max() < 1e-4
    assert np.array_equal(ids.numpy(), g["ids"])          # bit-exact ids


def test_encoder_causal_prefix(weights, gold):
    """Prefix causality of encode() (SURVEY.md section 8a-E note i): tokens of x[:N] equal the
    first N tokens of x."""
    g = gold("encoder_40f")
    wav = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: 24 * 2048][None]
    with torch.no_grad():
        ids, _ = E.encode(wav, weights["tok"])
    assert np.array_equal(ids.numpy()[0, 0], g["ids"][0, 0, :24])


def test_encoder_streaming_window(weights, gold):
    g = gold("encoder_window128")
    live = int(g["live_frames"])
    win = torch.zeros(1, 128 * 2048)
    win[:, -live * 2048:] = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: live * 2048]
    with torch.no_grad():
        ids, _ = E.encode(win, weights["tok"])
    assert np.array_equal(ids.numpy(), g["ids"])


def test_vocoder_20_frames(weights, gold):
    g = gold("vocoder_20f")
    codes = torch.from_numpy(g["codes"])
    with torch.no_grad():
        z = V.quantizer_decode(codes, weights["voc_folded"])
        wave = V.head(z, weights["voc_folded"])
    assert np.abs(z[0, :, -8:].numpy() - g["z_tail"]).max() < 1e-4
    mse = float(((wave[0, 0].numpy() - g["wave"]) ** 2).mean())
    assert mse < 1e-10, mse                                  # fp32 waveform MSE tolerance


def test_ar_streaming_codes(weights, gold, tape):
    g = gold("ar_stream")
    ar = DualAR(weights["ar"], tape(int(g["tape_seed"])))
    style, timbre = synth.synth_speaker(int(g["spk_seed"]))
    src = torch.from_numpy(g["src_content"])
    with torch.no_grad():
        ar.set_delay(int(g["delay"]))
        ar.prefill_prompt(torch.from_numpy(g["ref_content"]), torch.from_numpy(g["ref_audio"]), style, timbre)
        ar.prefill_src_condition4delay(src[:, :2])
        for i, t in enumerate(range(2, 16)):
            codes, pos = ar.decode_one(src[:, t:t + 1])
            assert np.array_equal(codes.numpy(), g["codes"][i]), (i, codes.T, g["codes"][i].T)
            assert int(pos) == int(g["pos"][i])


def test_ar_logits(weights, gold, tape):
    g = gold("ar_logits")
    s = gold("ar_stream")
    ar = DualAR(weights["ar"], tape(int(s["tape_seed"])))
    style, timbre = synth.synth_speaker(int(s["spk_seed"]))
    src = torch.from_numpy(s["src_content"])
    with torch.no_grad():
        ar.set_delay(2)
        ar.prefill_prompt(torch.from_numpy(s["ref_content"]), torch.from_numpy(s["ref_audio"]), style, timbre)
        assert np.abs(ar.last_hidden.numpy() - g["prefill_hidden"]).max() < 2e-4
        assert np.abs(ar.last_logits[0].numpy() - g["prefill_logits"]).max() < 2e-4
        ar.prefill_src_condition4delay(src[:, :2])
        codes, _ = ar.decode_one(src[:, 2:3])
    assert np.abs(ar.last_hidden.numpy() - g["hidden"]).max() < 2e-4
    assert np.abs(ar.last_logits[0].numpy() - g["slow_logits"]).max() < 2e-4
    assert np.abs(torch.stack(ar.last_logits[1:]).numpy() - g["fast_logits"]).max() < 2e-4
    assert np.array_equal(codes.numpy(), g["codes"])


def test_ar_offline_generate(weights, gold, tape):
    g = gold("ar_generate")
    s = gold("ar_stream")
    ar = DualAR(weights["ar"], tape(int(s["tape_seed"])))
    style, timbre = synth.synth_speaker(int(s["spk_seed"]))
    with torch.no_grad():
        ar.set_delay(2)
        out = ar.generate(torch.from_numpy(s["ref_content"]), torch.from_numpy(s["ref_audio"]),
                          torch.from_numpy(s["src_content"])[:, :10], style, timbre)
    assert np.array_equal(out.numpy(), g["codes"])


def test_ar_offline_generate_with_sampling_kwargs(weights, gold, tape):
    """`generate(..., temperature=0.9, top_p=0.85)` against the unmodified reference (tests/golden/ar_generate_kwargs.npz,
    oracle/make_golden_generate_kwargs.py): the first frame is sampled with the DEFAULT arguments
    (dual_ar_stream.py:723), the later ones with the caller's."""
    g, s = gold("ar_generate_kwargs"), gold("ar_stream")
    ar = DualAR(weights["ar"], tape(int(g["tape_seed"])))
    style, timbre = synth.synth_speaker(int(s["spk_seed"]))
    with torch.no_grad():
        ar.set_delay(2)
        out = ar.generate(torch.from_numpy(s["ref_content"]), torch.from_numpy(s["ref_audio"]),
                          torch.from_numpy(s["src_content"])[:, : int(g["n_src"])], style, timbre,
                          temperature=float(g["temperature"]), top_p=float(g["top_p"]))
    assert np.array_equal(out.numpy(), g["codes"])
    assert np.array_equal(out.numpy()[:, :, 0], gold("ar_generate")["codes"][:, :, 0])      # frame 0: defaults


def _run_stream(weights, g, tape):
    so = StreamOracle(weights["ar"], weights["tok"], weights["voc_folded"], tape(int(g["tape_seed"])))
    n_ref, n_chunks = int(g["n_ref"]), int(g["n_chunks"])
    style, timbre = synth.synth_speaker(int(g["ref_seed"]))
    ref_wave = synth.synth_audio_44k(int(g["ref_seed"]), 3.5)[: n_ref * 2048][None]
    gen = torch.Generator().manual_seed(int(g["codes_seed"]))
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    with torch.no_grad():
        ref_content = E.encode(ref_wave, weights["tok"])[0].squeeze(0)
        assert np.array_equal(ref_content.numpy(), g["ref_content"])
        so.prefill_prompt(ref_audio, ref_content, style, timbre, max_prompt_frames=256, delay=int(g["delay"]))
        so.setup_stream_caches(int(g["encode_window_frames"]), int(g["decode_window_frames"]),
                               int(g["max_seq_frames"]), int(g["buffer_frames"]), 1)
        src = synth.synth_audio_44k(int(g["src_seed"]), 1.5)[: n_chunks * 2048].view(n_chunks, 2048)
        waves = [so.process_one_chunk(src[i][None]) for i in range(n_chunks)]
    return so, torch.cat(waves, dim=-1)


def test_stream_reprompt_loop(weights, gold, tape):
    """Whole loop, small windows, re-prompt (infer_arvc.py:547-564) firing every few frames."""
    g = gold("stream_reprompt")
    so, wave = _run_stream(weights, g, tape)
    assert np.array_equal(so.src_content_codes.numpy(), g["src_content"])
    assert np.array_equal(so.pred_codes.numpy(), g["pred_codes"])
    assert float(((wave[0].numpy() - g["wave"]) ** 2).mean()) < 1e-10


def test_stream_default_loop(weights, gold, tape):
    """Whole loop at the CLI defaults (encode window 128, decode window 64)."""
    g = gold("stream_default")
    so, wave = _run_stream(weights, g, tape)
    assert np.array_equal(so.src_content_codes.numpy(), g["src_content"])
    assert np.array_equal(so.pred_codes.numpy(), g["pred_codes"])
    assert float(((wave[0].numpy() - g["wave"]) ** 2).mean()) < 1e-10


def test_vocoder_encode_oracle_vs_reference(gold, weights):
    """`wav2target_fn` (reference wave -> codec ids of the prompt): the restatement of FireflyArchitecture.encode +
    the FSQ index arithmetic against the reference's own vocoder object (tests/golden/vocoder_encode.npz, written by
    oracle/make_golden_vocenc.py): ids bit-exact."""
    from oracle import vocoder as V
    from streamvoiceanon_b200 import synth
    g = gold("vocoder_encode")
    for n in "ab":
        wav = synth.synth_audio_44k(int(g[f"seed_{n}"]), 3.0)[: int(g[f"frames_{n}"]) * 2048][None]
        with torch.no_grad():
            codes = V.wav2codes(wav, weights["voc_enc"])
        assert codes.dtype == torch.int32
        assert np.array_equal(codes.numpy(), g[f"codes_{n}"]), n


def test_noise_mixing_oracle_vs_reference(gold):
    """the anonymisation mix of `InferenceWrapper.apply_noise_mixing` (infer_arvc.py:228-232)
    restated in oracle/prompt.py against outputs of the unmodified reference method on the same draws
    (tests/golden/noise_mix.npz, oracle/make_golden_noise_mix.py).  fp32, tolerance 1e-6 absolute (statistics are
    accumulated in a different order)."""
    from oracle import prompt as P
    g = gold("noise_mix")
    for n in g["names"]:
        y = P.apply_noise_mixing(g[f"x_{n}"], float(g[f"alpha_{n}"]), g[f"noise_{n}"])
        assert y.shape == g[f"y_{n}"].shape and y.dtype == np.float32
        assert np.abs(y - g[f"y_{n}"]).max() < 1e-6, n
    # alpha = 1 returns the input unchanged, alpha = 0 pure (re-scaled) noise
    assert np.array_equal(P.apply_noise_mixing(g["x_timbre_a1"], 1.0, g["noise_timbre_a1"]), g["x_timbre_a1"])


def test_style_vector_oracle_vs_reference(gold):
    """SURVEY section 8f-3, style branch (groundwork: oracle only, the CUDA port of this row is not built): kaldi fbank
    -> time-mean removal -> CAMPPlus restated in oracle/speaker.py against the reference's own
    `calculate_style_vec` + torchaudio fbank (tests/golden/style_vec.npz, oracle/make_golden_style.py).  fp32;
    tolerance 1e-4 on the log-mel features, 1e-5 on the embedding (summation order of the host BLAS)."""
    from oracle import speaker as S
    from streamvoiceanon_b200 import synth
    g = gold("style_vec")
    sd = synth.make_campplus_state_dict(int(g["weight_seed"]))
    a = synth.synth_audio_16k(int(g["seed_a"]), float(g["sec_a"]))[None]
    b = synth.synth_audio_16k(int(g["seed_b"]), float(g["sec_b"]))[None]
    fb = S.kaldi_fbank(a)
    assert tuple(fb.shape) == g["fbank_a"].shape == (1 + (a.shape[1] - 400) // 160, 80)
    assert np.abs(fb.numpy() - g["fbank_a"]).max() < 1e-4
    with torch.no_grad():
        sa = S.calculate_style_vec(a, torch.LongTensor([a.shape[1]]), sd)
        batch = torch.zeros(2, a.shape[1])
        batch[0], batch[1, : b.shape[1]] = a[0], b[0]
        sb = S.calculate_style_vec(batch, torch.from_numpy(g["batch_lens"]), sd)
    assert np.abs(sa.numpy() - g["style_a"]).max() < 1e-5
    assert np.abs(sb.numpy() - g["style_batch"]).max() < 1e-5
    # ragged row: padding with the row minimum must not leak into the masked statistics of the short row's ... long row
    assert np.abs(sb.numpy()[0] - sa.numpy()[0]).max() < 1e-5
    assert S.kaldi_fbank(a[:, :399]).shape[0] == 0               # shorter than one 25 ms frame: no frames


def test_timbre_latent_oracle_vs_reference(gold):
    """SURVEY section 8f-3, timbre branch (groundwork: oracle only): slaney mel -> ECAPA-TDNN trunk -> PerceiverResampler ->
    FSQ restated in oracle/speaker.py against the reference's own `calculate_timbre_latent` / `tokenize_wav` +
    torchaudio MelSpectrogram (tests/golden/timbre_latent.npz, oracle/make_golden_style.py).  FSQ indices exact and
    latents to 1e-5 wherever the pre-rounding coordinate is further than 1e-3 from a rounding boundary."""
    from oracle import speaker as S
    from streamvoiceanon_b200 import synth
    g = gold("timbre_latent")
    sd = synth.make_timbre_encoder_state_dict(int(g["weight_seed"]))
    a = synth.synth_audio_16k(int(g["seed_a"]), float(g["sec_a"]))[None]
    b = synth.synth_audio_16k(int(g["seed_b"]), float(g["sec_b"]))[None]
    mel = S.timbre_mel(a)
    assert tuple(mel.shape) == g["mel_a"].shape == (1, a.shape[1] // 320 + 1, 128)
    assert np.abs(mel.numpy() - g["mel_a"]).max() < 1e-4
    batch = torch.zeros(2, a.shape[1])
    batch[0], batch[1, : b.shape[1]] = a[0], b[0]
    with torch.no_grad():
        runs = ((S.calculate_timbre_latent(a, torch.LongTensor([a.shape[1]]), sd), "a"),
                (S.calculate_timbre_latent(batch, torch.from_numpy(g["batch_lens"]), sd), "batch"))
    for (zq, idx, bounded), name in runs:
        assert tuple(zq.shape) == g[f"timbre_{name}"].shape and idx.dtype == torch.int32
        frac = (bounded - bounded.floor() - 0.5).abs()                       # distance to the .5 rounding boundary
        safe = (frac > 1e-3).all(dim=-1).numpy()                             # [B, 32] tokens that cannot flip
        assert safe.mean() > 0.9
        assert np.array_equal(idx.numpy()[safe], g[f"indices_{name}"][:, 0][safe]), name
        assert np.abs(zq.numpy() - g[f"timbre_{name}"])[safe].max() < 1e-5, name
    # the mask keeps padded frames of the short row out of the attention: row 0 of the batch equals the single run
    assert np.array_equal(runs[1][0][1].numpy()[0], runs[0][0][1].numpy()[0])


def test_calculate_prompt_oracle_vs_reference(gold):
    """BASELINE config 5's prompt: the UNMODIFIED `InferenceWrapper.calculate_prompt` (infer_arvc.py:382-441) on three
    concatenated references with alpha = 0.7 (tests/golden/prompt_config5.npz, oracle/make_golden_prompt.py) against the
    chained restatement oracle/prompt.py::calculate_prompt (resample -> both speaker encoders -> noise mix in the
    reference's draw order -> vocoder encoder ids -> content ids).  Ids exact; embeddings to 1e-5."""
    import torchaudio
    from oracle import prompt as P
    from streamvoiceanon_b200 import synth
    g = gold("prompt_config5")
    ws = int(g["weight_seed"])
    refs = [synth.synth_audio_44k(int(s), float(g["ref_seconds"]))[None] for s in g["ref_seeds"]]
    with torch.no_grad():
        codes, content, style, timbre, ref = P.calculate_prompt(
            refs, float(g["alpha"]), g["noise_style"], g["noise_timbre"], synth.make_campplus_state_dict(ws),
            synth.make_timbre_encoder_state_dict(ws), synth.make_tokenizer_state_dict(ws),
            synth.make_vocoder_encoder_state_dict(ws))
    assert ref.shape[-1] == int(g["n_samples"])
    assert codes.dtype == torch.int32 and np.array_equal(codes.numpy(), g["ref_audio_codes"])
    assert np.array_equal(content.numpy(), g["ref_content_codes"])
    assert np.abs(style.numpy() - g["style_vectors"]).max() < 1e-5
    assert np.abs(timbre.numpy() - g["timbre_latents"]).max() < 1e-5
    # the 16 kHz step is torchaudio.functional.resample, whose filter bank is built in the waveform's dtype
    want = torchaudio.functional.resample(ref, 44100, 16000)
    assert float((P.resample(ref, 44100, 16000) - want).abs().max()) < 1e-6


def test_stream_config5_loop_from_reference_waves(weights, gold, tape):
    """BASELINE config 5 end to end on the CPU: three reference WAVES -> oracle calculate_prompt (speaker encoders, noise
    mix, codec and content ids) -> prompt truncated to 48 frames -> the streaming loop with two-frame chunks and the
    re-prompt path firing, against the unmodified reference's `prefill_prompt` + `process_one_chunk`
    (tests/golden/stream_config5.npz, oracle/make_golden_prompt.py): ids exact, waveform MSE < 1e-10."""
    from oracle import prompt as P
    g, gp = gold("stream_config5"), gold("prompt_config5")
    ws = int(g["weight_seed"])
    refs = [synth.synth_audio_44k(int(s), float(gp["ref_seconds"]))[None] for s in gp["ref_seeds"]]
    with torch.no_grad():
        codes, content, style, timbre, _ = P.calculate_prompt(
            refs, float(g["alpha"]), gp["noise_style"], gp["noise_timbre"], synth.make_campplus_state_dict(ws),
            synth.make_timbre_encoder_state_dict(ws), weights["tok"], weights["voc_enc"])
        so = StreamOracle(weights["ar"], weights["tok"], weights["voc_folded"], tape(int(g["tape_seed"])))
        so.prefill_prompt(codes, content, style, timbre, max_prompt_frames=int(g["max_prompt_frames"]), delay=int(g["delay"]))
        chunk = int(g["decode_chunk_frames"])
        so.setup_stream_caches(int(g["encode_window_frames"]), int(g["decode_window_frames"]), int(g["max_seq_frames"]),
                               int(g["buffer_frames"]), chunk)
        n = int(g["n_chunks"])
        src = synth.synth_audio_44k(int(g["src_seed"]), 1.5)[: n * chunk * 2048].view(n, chunk * 2048)
        wave = torch.cat([so.process_one_chunk(src[i][None]) for i in range(n)], dim=-1)
    assert np.array_equal(so.src_content_codes.numpy(), g["src_content"])
    assert np.array_equal(so.pred_codes.numpy(), g["pred_codes"])
    assert float(((wave[0].numpy() - g["wave"]) ** 2).mean()) < 1e-10


def _infer_inputs(g):
    src = synth.synth_audio_44k(int(g["src_seed"]), float(g["src_seconds"]))[None]
    refs = [synth.synth_audio_44k(int(s), float(g["ref_seconds"]))[None] for s in g["ref_seeds"]]
    return src, refs


@pytest.mark.parametrize("collate,key", [("concat_mel", "wave"), ("avg", "wave_avg")])
def test_offline_infer_from_files_oracle_vs_reference(weights, gold, tape, collate, key):
    """BASELINE config 1 as a user runs it: the UNMODIFIED `InferenceWrapper.infer(src.wav, [ref_a.wav, ref_b.wav],
    delay=2, alpha=0.7)` (infer_arvc.py:261-380; tests/golden/infer_config1.npz, oracle/make_golden_infer.py) against the
    chained oracle: calculate_prompt (both speaker encoders, mix with the recorded draws, codec + content ids) ->
    tokenizer on the source -> offline `generate` -> code2wav, with both speaker-embedding collations ("avg": each
    reference's embeddings on their own, averaged, :282-307).  Waveform MSE < 1e-10 (ids are exact or it would not be)."""
    from oracle import prompt as P
    from oracle import vocoder as V
    g = gold("infer_config1")
    ws = int(g["weight_seed"])
    src, refs = _infer_inputs(g)
    with torch.no_grad():
        codes, content, style, timbre, _ = P.calculate_prompt(
            refs, float(g["alpha"]), g["noise_style"], g["noise_timbre"], synth.make_campplus_state_dict(ws),
            synth.make_timbre_encoder_state_dict(ws), weights["tok"], weights["voc_enc"], collate=collate)
        src_content = E.encode(src, weights["tok"])[0].squeeze(0)
        ar = DualAR(weights["ar"], tape(int(g["tape_seed"])))
        ar.set_delay(2)
        vc = ar.generate(content, codes, src_content, style, timbre)
        wave = V.code2wav(vc.long(), weights["voc_folded"]).squeeze()
    assert wave.shape == g[key].shape == (src.shape[1] // 2048 * 2048,)
    assert float(((wave.numpy() - g[key]) ** 2).mean()) < 1e-10


def test_stream_infer_from_files_oracle_vs_reference(weights, gold, tape):
    """BASELINE config 2 as a user runs it: the UNMODIFIED `InferenceWrapper.stream_infer(src.wav, ref.wav, chunk 1,
    delay 2)` (infer_arvc.py:598-676) -- file loading, prompt from the reference WAVE, left padding to whole chunks, the
    loop with the re-prompt firing -- against the chained oracle.  Ids exact, waveform MSE < 1e-10."""
    from oracle import prompt as P
    g = gold("infer_config1")
    ws = int(g["weight_seed"])
    src, refs = _infer_inputs(g)
    cfg = {k: int(g[f"stream_{k}"]) for k in ("encode_window_frames", "decode_window_frames", "max_prompt_frames",
                                              "max_seq_frames", "buffer_frames", "decode_chunk_frames", "delay")}
    with torch.no_grad():
        z192, z4096 = np.zeros((1, 192), np.float32), np.zeros((1, 32, 128), np.float32)
        codes, content, style, timbre, _ = P.calculate_prompt(                      # alpha = 1: the draws do not matter
            refs[:1], 1.0, z192, z4096, synth.make_campplus_state_dict(ws), synth.make_timbre_encoder_state_dict(ws),
            weights["tok"], weights["voc_enc"])
        so = StreamOracle(weights["ar"], weights["tok"], weights["voc_folded"], tape(int(g["tape_seed"])))
        so.prefill_prompt(codes, content, style, timbre, max_prompt_frames=cfg["max_prompt_frames"], delay=cfg["delay"])
        so.setup_stream_caches(cfg["encode_window_frames"], cfg["decode_window_frames"], cfg["max_seq_frames"],
                               cfg["buffer_frames"], cfg["decode_chunk_frames"])
        step = 2048 * cfg["decode_chunk_frames"]
        padded = torch.nn.functional.pad(src, (step - src.shape[1] % step, 0)).view(-1, step)
        wave = torch.cat([so.process_one_chunk(padded[i][None]) for i in range(padded.shape[0])], dim=-1)
    assert np.array_equal(so.src_content_codes.numpy(), g["stream_src_content"])
    assert np.array_equal(so.pred_codes.numpy(), g["stream_pred_codes"])
    assert float(((wave[0].numpy() - g["stream_wave"]) ** 2).mean()) < 1e-10


def test_speaker_encoders_full_size_oracle_vs_reference(gold):
    """Both speaker encoders at BASELINE config 5's full size (15 s = three 5 s references: 1498 fbank frames, 8 CAM
    segments, 751 mel frames) against the unmodified reference (tests/golden/speaker_full_15s.npz,
    oracle/make_golden_speaker_full.py)."""
    from oracle import speaker as S
    g = gold("speaker_full_15s")
    ws = int(g["weight_seed"])
    wave = torch.cat([synth.synth_audio_16k(int(s), float(g["seconds"])) for s in g["seeds"]])[None]
    lens = torch.LongTensor([wave.shape[1]])
    with torch.no_grad():
        style = S.calculate_style_vec(wave, lens, synth.make_campplus_state_dict(ws))
        zq, idx, bounded = S.calculate_timbre_latent(wave, lens, synth.make_timbre_encoder_state_dict(ws))
    assert np.abs(style.numpy() - g["style"]).max() < 1e-5
    safe = ((bounded - bounded.floor() - 0.5).abs() > 1e-3).all(dim=-1).numpy()
    assert safe.mean() > 0.9
    assert np.array_equal(idx.numpy()[safe], g["indices"][:, 0][safe])
    assert np.abs(zq.numpy() - g["timbre"])[safe].max() < 1e-5


def test_named_inputs_oracle_vs_reference(weights, gold, tape):
    """BASELINE configs 1-2 on the inputs BASELINE.json names (tests/golden/trump_0.wav -> azuma_0.wav; fixture
    tests/golden/config12_named.npz from the unmodified reference, oracle/make_golden_named.py): the oracle's prompt from
    the 153-frame reference wave (codec ids, content ids exact; speaker embeddings 1e-5) and the first 10 chunks of the
    CLI-default streaming loop (content ids and codec ids exact; waveform MSE < 1e-10).  Bounded so that the CPU suite
    stays short; the GPU tests run all 168 chunks and the offline call."""
    from pathlib import Path
    from scipy.io import wavfile
    from oracle import prompt as P
    GOLD = Path(__file__).resolve().parent / "golden"
    g = gold("config12_named")
    ws = int(g["weight_seed"])

    def load(name):
        rate, data = wavfile.read(str(GOLD / name))
        assert rate == 44100 and data.dtype == np.int16
        x = torch.from_numpy(data.astype(np.float32) / 32768.0)
        return (x.mean(dim=1) if x.dim() == 2 else x)[None]
    src, ref = load("trump_0.wav"), load("azuma_0.wav")
    assert src.shape[1] == 343483 and ref.shape[1] // 2048 == 153
    n = 10
    with torch.no_grad():
        z192, z4096 = np.zeros((1, 192), np.float32), np.zeros((1, 32, 128), np.float32)
        codes, content, style, timbre, _ = P.calculate_prompt([ref], 1.0, z192, z4096, synth.make_campplus_state_dict(ws),
                                                              synth.make_timbre_encoder_state_dict(ws), weights["tok"],
                                                              weights["voc_enc"])
        assert np.array_equal(content.numpy(), g["ref_content"])
        assert np.array_equal(codes.numpy(), g["ref_audio"])
        assert np.abs(style.numpy() - g["style"]).max() < 1e-5
        so = StreamOracle(weights["ar"], weights["tok"], weights["voc_folded"], tape(int(g["tape_seed"])))
        so.prefill_prompt(codes, content, style, timbre, max_prompt_frames=256, delay=2)
        so.setup_stream_caches(128, 64, 768, 32, 1)
        padded = torch.nn.functional.pad(src, (2048 - src.shape[1] % 2048, 0)).view(-1, 2048)
        assert padded.shape[0] == 168
        wave = torch.cat([so.process_one_chunk(padded[i][None]) for i in range(n)], dim=-1)
    assert np.array_equal(so.src_content_codes.numpy(), g["stream_src_content"][:, :n])
    assert np.array_equal(so.pred_codes.numpy(), g["stream_pred_codes"][:, :, : n - 2])
    assert float(((wave[0].numpy() - g["stream_wave"][: n * 2048]) ** 2).mean()) < 1e-10

"""The reference's UNMODIFIED caller over the engine ("configs 1, 2, 5 through the unmodified
`InferenceWrapper` with the engine swapped in behind L3"; north_star: "drops into infer_arvc.py unchanged").

`evaluations.infer_arvc.InferenceWrapper` -- the reference's own class, byte-compiled into oracle/_ref by
oracle/build_ref.py (or imported from /root/reference where that exists) -- is CONSTRUCTED by its own `__init__`
(:33-144: YAML -> `hydra.utils.instantiate` -> `load_state_dict` from checkpoint files -> `setup_caches` / `.to()` /
`.eval()` / `remove_parametrizations()`) from the drop-in YAML set of INTEGRATION.md section 2
(tools/make_dropin_configs.py: five `_target_`s pointed at the shims), then its own `infer` and `stream_infer` run from
.wav files: file loading, resampling and kaldi fbank in the caller's torch code, `calculate_prompt`, `prefill_prompt`,
`setup_stream_caches`, `process_one_chunk` (three shim calls per chunk between the caller's synchronisations), padding
rule, re-prompt.  Results against tests/golden/infer_config1.npz, which the same class produced with the reference's
own five torch modules on the CPU: ids exact, waveform MSE < 1e-8.

Test-side patches, none of them in the reference: the sampling noise tape is set on the AR shim (the reference arm of the
fixture patched `multinomial_sample_one_no_sync`), and `torch.randn_like` replays the two recorded draws of the
anonymisation mix (CUDA and CPU generators differ)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from streamvoiceanon_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]


@pytest.fixture(scope="module")
def caller(tmp_path_factory, gold):
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("neither /root/reference nor oracle/_ref (python -m oracle.build_ref in the build container) is present")
    sys.path.insert(0, str(ROOT / "tools"))
    import make_dropin_configs
    ref_harness._paths()                                   # import shims (hydra, omegaconf, librosa, ...) + the reference root
    work = tmp_path_factory.mktemp("dropin")
    seed = int(gold("infer_config1")["weight_seed"])
    ck = work / "pretrained_checkpoints"
    ck.mkdir()
    top = make_dropin_configs.make(ref_harness.REF_ROOT, work, checkpoints=ck)
    import yaml
    cfg = yaml.safe_load(open(top))
    voc_sd = dict(synth.make_vocoder_state_dict(seed))
    voc_sd.update(synth.make_vocoder_encoder_state_dict(seed))
    torch.save(synth.make_tokenizer_state_dict(seed), cfg["speech_tokenizer"]["checkpoint_path"])
    torch.save(voc_sd, cfg["firefly"]["checkpoint_path"])
    torch.save(synth.make_campplus_state_dict(seed), cfg["style_encoder"]["checkpoint_path"])
    torch.save(synth.make_timbre_encoder_state_dict(seed), cfg["timbre_encoder"]["checkpoint_path"])
    ar_ckpt = ck / "dual_ar_delay_0_8.pth"
    torch.save(synth.make_ar_state_dict(seed), ar_ckpt)
    cwd = os.getcwd()
    os.chdir(work)                                         # the reference resolves its YAML / checkpoint paths against the cwd
    try:
        from evaluations.infer_arvc import InferenceWrapper
        iw = InferenceWrapper(str(top), str(ar_ckpt), compile_encoder=False, compile_decoder=False, compile_ar=True, fp16=True)
    finally:
        os.chdir(cwd)
    assert type(iw).__module__ == "evaluations.infer_arvc"
    assert type(iw.model).__module__ == "streamvoiceanon_b200.arvc_wrapper"
    assert type(iw.speech_tokenizer).__module__ == "streamvoiceanon_b200.firefly"
    assert type(iw.style_encoder).__module__ == "streamvoiceanon_b200.speaker"
    return iw, work


class _ReplayRandn:
    """torch.randn_like stand-in that hands out recorded draws in order (style first, then timbre, :419-421)."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.real = torch.randn_like

    def __enter__(self):
        def fake(t, *a, **k):
            if self.draws and tuple(self.draws[0].shape) == tuple(t.shape):
                return self.draws.pop(0).to(device=t.device, dtype=t.dtype)
            return self.real(t, *a, **k)
        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self.real


def _wavs(work, g):
    src = work / "src.wav"
    wavfile.write(src, 44100, synth.synth_audio_44k(int(g["src_seed"]), float(g["src_seconds"])).numpy())
    refs = []
    for s in g["ref_seeds"]:
        p = work / f"ref{int(s)}.wav"
        wavfile.write(p, 44100, synth.synth_audio_44k(int(s), float(g["ref_seconds"])).numpy())
        refs.append(str(p))
    return str(src), refs


def test_unmodified_infer_config1(caller, gold, tape):
    """BASELINE config 1's call: `InferenceWrapper.infer(src.wav, [a.wav, b.wav], delay=2, alpha=0.7)` (:261-380)."""
    iw, work = caller
    g = gold("infer_config1")
    src, refs = _wavs(work, g)
    for collate, key in (("concat_mel", "wave"), ("avg", "wave_avg")):
        iw.model.set_noise_fn(tape(int(g["tape_seed"])), 0)
        with _ReplayRandn([torch.from_numpy(g["noise_style"]), torch.from_numpy(g["noise_timbre"])]):
            wave = iw.infer(src, refs, delay=2, alpha=float(g["alpha"]), spk_emb_collate_type=collate, save_result=False)
        wave = np.asarray(wave, dtype=np.float32).reshape(-1)
        assert wave.shape == g[key].shape, collate
        mse = float(((wave - g[key]) ** 2).mean())
        assert mse < 1e-8, (collate, mse)


def test_unmodified_stream_infer_config2(caller, gold, tape):
    """BASELINE config 2's call: `InferenceWrapper.stream_infer(src.wav, a.wav, decode_chunk_frames=1, delay=2)` (:598-676)
    -- the reference's own `process_one_chunk` loop (with a `pitch_shift` keyword once, as the GUI passes it)."""
    iw, work = caller
    g = gold("infer_config1")
    src, refs = _wavs(work, g)
    cfg = {k: int(g[f"stream_{k}"]) for k in ("encode_window_frames", "decode_window_frames", "max_prompt_frames",
                                              "max_seq_frames", "buffer_frames", "decode_chunk_frames", "delay")}
    iw.model.set_noise_fn(tape(int(g["tape_seed"])), 0)
    wave = iw.stream_infer(src, refs[0], save_result=False, alpha=1.0, **cfg)
    assert np.array_equal(iw.src_content_codes.cpu().numpy(), g["stream_src_content"])
    assert np.array_equal(iw.pred_codes.cpu().numpy(), g["stream_pred_codes"])
    wave = np.asarray(wave, dtype=np.float32).reshape(-1)
    assert float(((wave - g["stream_wave"]) ** 2).mean()) < 1e-8
    out = iw.process_one_chunk(torch.zeros(1, 2048, device=iw.device), pitch_shift=0.0)     # real-time-gui.py's call shape
    assert tuple(out.shape) == (1, 2048)


def test_unmodified_caller_named_inputs_config1_config2(caller, gold, tape):
    """BASELINE configs 1 and 2 on the inputs BASELINE.json names (tests/golden/trump_0.wav -> azuma_0.wav: 167 source
    frames, 153 prompt frames), the reference's own calls with the CLI-default windows, against the unmodified
    reference's CPU run with its own modules (tests/golden/config12_named.npz, oracle/make_golden_named.py)."""
    iw, _ = caller
    g = gold("config12_named")
    src, ref = str(ROOT / "tests" / "golden" / "trump_0.wav"), str(ROOT / "tests" / "golden" / "azuma_0.wav")
    iw.model.set_noise_fn(tape(int(g["tape_seed"])), 0)
    wave = np.asarray(iw.infer(src, ref, delay=2, save_result=False), dtype=np.float32).reshape(-1)
    assert wave.shape == g["wave"].shape == (167 * 2048,)
    assert float(((wave - g["wave"]) ** 2).mean()) < 1e-8
    iw.model.set_noise_fn(tape(int(g["tape_seed"])), 0)
    stream = np.asarray(iw.stream_infer(src, ref, decode_chunk_frames=1, delay=2, save_result=False), dtype=np.float32).reshape(-1)
    assert tuple(iw.ref_content_codes.shape) == (1, 153)
    assert np.array_equal(iw.ref_content_codes.cpu().numpy(), g["ref_content"])
    assert np.array_equal(iw.ref_audio_codes.cpu().numpy(), g["ref_audio"])
    assert np.array_equal(iw.src_content_codes.cpu().numpy(), g["stream_src_content"])
    assert np.array_equal(iw.pred_codes.cpu().numpy(), g["stream_pred_codes"])
    assert stream.shape == g["stream_wave"].shape == (168 * 2048,)
    assert float(((stream - g["stream_wave"]) ** 2).mean()) < 1e-8

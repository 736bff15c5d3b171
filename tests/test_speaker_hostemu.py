"""Host orchestration of the speaker-encoder port (streamvoiceanon_b200/csrc/speaker.hpp) against
the reference-generated fixtures, on a machine WITHOUT a GPU.

speaker.hpp is written against a backend interface; tests/hostemu/speaker_host.cpp compiles the same source with g++:
functors run as loops and the engine's GEMM is a three-loop restatement of the GemmParams contract.  What this pins:
buffer shapes and zero margins, weight repacking, GEMM descriptors (overlapping rows, row-offset taps, column slices of
the concat buffers), functor arguments and the derived buffers built by streamvoiceanon_b200/speaker.py.  What it
cannot pin: the CUDA launch of the functors and the real GEMM kernels -- tests/test_zz_gpu_speaker.py does that on the
B200.  The host build is test infrastructure: the product never loads it (tests/test_cabi.py)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from streamvoiceanon_b200 import synth

HERE = Path(__file__).resolve().parent / "hostemu"
SRC = HERE / "speaker_host.cpp"
HPP = HERE.parent.parent / "streamvoiceanon_b200" / "csrc" / "speaker.hpp"
SO = HERE / "_build" / "libspeaker_host.so"


@pytest.fixture(scope="module")
def emu():
    SO.parent.mkdir(exist_ok=True)
    if not SO.exists() or SO.stat().st_mtime < max(SRC.stat().st_mtime, HPP.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", str(SRC),
                        "-o", str(SO)], check=True)
    lib = C.CDLL(str(SO))
    lib.hostemu_last_error.restype = C.c_char_p
    from streamvoiceanon_b200 import speaker as SP

    def load(model, sd):
        for k, v in sd.items():
            if not torch.is_tensor(v) or not v.is_floating_point() or v.dim() == 0:
                continue
            t = v.detach().float().contiguous()
            shape = (C.c_longlong * t.dim())(*t.shape)
            assert lib.hostemu_load_tensor(model, k.encode(), C.c_void_p(t.data_ptr()), t.dim(), shape) == 0
        assert lib.hostemu_finalize(model) == 0, lib.hostemu_last_error().decode()

    seed = int(np.load(HERE.parent / "golden" / "style_vec.npz")["weight_seed"])
    load(0, {**synth.make_campplus_state_dict(seed), **SP.style_derived_buffers()})
    load(1, {**synth.make_timbre_encoder_state_dict(seed), **SP.timbre_derived_buffers()})
    return lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def test_style_branch_host_orchestration_vs_reference(emu, gold):
    """kaldi fbank to 1e-4 and the style vector to 1e-4 against the unmodified reference (`calculate_style_vec`,
    tests/golden/style_vec.npz); the ragged batch row goes through fbank + CAMPPlus with the reference's padding rule."""
    g = gold("style_vec")
    a = synth.synth_audio_16k(int(g["seed_a"]), float(g["sec_a"])).contiguous()
    b = synth.synth_audio_16k(int(g["seed_b"]), float(g["sec_b"])).contiguous()
    T = 1 + (a.numel() - 400) // 160
    feat = torch.empty(T, 80)
    assert emu.hostemu_kaldi_fbank(_p(a), C.c_longlong(a.numel()), _p(feat)) == 0, emu.hostemu_last_error().decode()
    assert np.abs(feat.numpy() - g["fbank_a"]).max() < 1e-4
    out = torch.empty(192)
    counts = (C.c_longlong * 3)()
    assert emu.hostemu_style_vector(_p(a), C.c_longlong(a.numel()), _p(out), counts) == 0, emu.hostemu_last_error().decode()
    assert np.isfinite(out.numpy()).all()
    assert np.abs(out.numpy() - g["style_a"][0]).max() < 1e-4
    assert counts[0] == 1 + 52 * 2 + 3                           # TDNN, 52 x (bottleneck, local conv), 3 transits
    # the short row of the reference's ragged batch: its own fbank minus its mean, padded with its minimum to the long
    # row's length, lens = frames // 2
    Tb = 1 + (b.numel() - 400) // 160
    fb = torch.empty(Tb, 80)
    assert emu.hostemu_kaldi_fbank(_p(b), C.c_longlong(b.numel()), _p(fb)) == 0
    fb = fb - fb.mean(dim=0, keepdim=True)
    padded = torch.nn.functional.pad(fb, (0, 0, 0, T - Tb), value=float(fb.min())).contiguous()
    assert emu.hostemu_campplus(_p(padded), C.c_longlong(T), Tb // 2, _p(out)) == 0, emu.hostemu_last_error().decode()
    assert np.abs(out.numpy() - g["style_batch"][1]).max() < 1e-4


def test_timbre_branch_host_orchestration_vs_reference(emu, gold):
    """Timbre latents and FSQ indices against the unmodified reference (`calculate_timbre_latent`,
    tests/golden/timbre_latent.npz): indices exact and latents to 1e-4 wherever the FSQ input is further than 1e-3 from a
    rounding boundary; single utterance and the zero-padded short row of the reference's batch (mask = wave_len // 320)."""
    from oracle import speaker as S
    g = gold("timbre_latent")
    a = synth.synth_audio_16k(int(g["seed_a"]), float(g["sec_a"])).contiguous()
    b = synth.synth_audio_16k(int(g["seed_b"]), float(g["sec_b"]))
    row_b = torch.zeros(a.numel())
    row_b[: b.numel()] = b
    for wave, wave_len, want, want_idx in ((a, a.numel(), g["timbre_a"][0], g["indices_a"][0, 0]),
                                           (row_b, b.numel(), g["timbre_batch"][1], g["indices_batch"][1, 0])):
        out, idx, z = torch.empty(32, 128), torch.empty(32, dtype=torch.int32), torch.empty(32, 6)
        counts = (C.c_longlong * 3)()
        rc = emu.hostemu_timbre_latent(_p(wave), C.c_longlong(wave.numel()), C.c_longlong(wave_len), _p(out), _p(idx), _p(z), counts)
        assert rc == 0, emu.hostemu_last_error().decode()
        assert np.isfinite(out.numpy()).all()
        _, _, bounded = S.fsq4_quantize(z)
        safe = ((bounded - bounded.floor() - 0.5).abs() > 1e-3).all(dim=-1).numpy()
        assert safe.mean() > 0.9
        assert np.array_equal(idx.numpy()[safe], want_idx[safe])
        assert np.abs(out.numpy() - want)[safe].max() < 1e-4
        assert counts[0] == 1 + 3 * 9 + 1 + 1 + 2                # layer1, 3 x (in, 7 res2, out), cat conv, context, 2 x kv


def test_speaker_functions_reject_short_waves(emu):
    out = torch.empty(192)
    w = torch.zeros(700)
    assert emu.hostemu_style_vector(_p(w), C.c_longlong(700), _p(out), None) == 1
    assert b"shorter" in emu.hostemu_last_error()
    lat = torch.empty(32, 128)
    assert emu.hostemu_timbre_latent(_p(w), C.c_longlong(700), C.c_longlong(700), _p(lat), None, None, None) == 1


def _style(emu, wave):
    out = torch.empty(192)
    assert emu.hostemu_style_vector(_p(wave), C.c_longlong(wave.numel()), _p(out), None) == 0, emu.hostemu_last_error().decode()
    return out


def _timbre(emu, wave, wave_len=None):
    out, idx, z = torch.empty(32, 128), torch.empty(32, dtype=torch.int32), torch.empty(32, 6)
    n = wave.numel()
    rc = emu.hostemu_timbre_latent(_p(wave), C.c_longlong(n), C.c_longlong(n if wave_len is None else wave_len), _p(out), _p(idx),
                                   _p(z), None)
    assert rc == 0, emu.hostemu_last_error().decode()
    return out, idx, z


def test_config5_prompt_embeddings_host_orchestration(emu, gold):
    """The 4.8 s three-reference concatenation of BASELINE config 5: both embeddings through the host build, mixed with
    the reference's recorded draws (oracle mix), against the outputs of the unmodified `calculate_prompt`."""
    from oracle import prompt as P
    gp = gold("prompt_config5")
    refs = [synth.synth_audio_44k(int(s), float(gp["ref_seconds"]))[None] for s in gp["ref_seeds"]]
    ref16 = P.resample(torch.cat(refs, dim=-1), 44100, 16000)[0].contiguous()
    alpha = float(gp["alpha"])
    sv = P.apply_noise_mixing(_style(emu, ref16).numpy()[None], alpha, gp["noise_style"])
    tl = P.apply_noise_mixing(_timbre(emu, ref16)[0].numpy()[None], alpha, gp["noise_timbre"])
    assert np.abs(sv - gp["style_vectors"]).max() < 1e-4
    assert np.abs(tl - gp["timbre_latents"]).max() < 1e-4


@pytest.mark.parametrize("n", [880, 1024, 16000 + 37, 33333])
def test_edge_lengths_host_orchestration_vs_oracle(emu, gold, n):
    """Shortest accepted waves (4 fbank frames: every dilated conv reads mostly margin rows; 1024 samples: 4 mel frames),
    a length that is no multiple of either hop, and one with several CAM segments whose last one is short -- against the
    CPU oracle (itself pinned to the reference)."""
    from oracle import speaker as S
    seed = int(gold("style_vec")["weight_seed"])
    wave = synth.synth_audio_16k(5300 + n % 7, 2.5)[:n].contiguous()
    lens = torch.LongTensor([n])
    with torch.no_grad():
        want_s = S.calculate_style_vec(wave[None], lens, synth.make_campplus_state_dict(seed))
        want_t, want_idx, bounded = S.calculate_timbre_latent(wave[None], lens, synth.make_timbre_encoder_state_dict(seed))
    assert np.abs(_style(emu, wave).numpy() - want_s[0].numpy()).max() < 1e-4
    if n >= 1024:
        out, idx, _ = _timbre(emu, wave)
        safe = ((bounded[0] - bounded[0].floor() - 0.5).abs() > 1e-3).all(dim=-1).numpy()
        assert np.array_equal(idx.numpy()[safe], want_idx[0].numpy()[safe])
        assert np.abs(out.numpy() - want_t[0].numpy())[safe].max() < 1e-4


def test_tensor_core_product_arithmetic_keeps_the_embeddings(emu, gold, monkeypatch):
    """The engine runs these GEMMs on the tensor cores as 3xTF32 products (hi*hi + hi*lo + lo*hi, gemm_tc.cu).  With the
    host GEMM restating that arithmetic the style vector stays within 1e-5 of the reference and the timbre latents stay
    exact -- the GPU tolerances (2e-4 / 1e-4) are not there to absorb the GEMM arithmetic."""
    monkeypatch.setenv("HOSTEMU_TF32X3", "1")
    gs, gt = gold("style_vec"), gold("timbre_latent")
    a = synth.synth_audio_16k(int(gs["seed_a"]), float(gs["sec_a"])).contiguous()
    assert np.abs(_style(emu, a).numpy() - gs["style_a"][0]).max() < 1e-5
    out, idx, _ = _timbre(emu, a)
    assert np.array_equal(idx.numpy(), gt["indices_a"][0, 0])
    assert np.abs(out.numpy() - gt["timbre_a"][0]).max() < 1e-5


def test_derived_buffers_are_torchaudios():
    """The four buffers the shim uploads next to the checkpoints are the ones torchaudio builds for the reference:
    kaldi mel banks + povey window (torchaudio.compliance.kaldi), slaney filterbank + the window torch.stft pads to n_fft."""
    import torchaudio
    from torchaudio.compliance import kaldi
    from streamvoiceanon_b200 import speaker as SP
    banks, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    mine = SP.kaldi_mel_banks()
    assert tuple(mine.shape) == (80, 257) and torch.equal(mine[:, :256], banks) and float(mine[:, 256].abs().max()) == 0.0
    povey = kaldi._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32)
    assert torch.equal(SP.povey_window(), povey)
    fb = torchaudio.functional.melscale_fbanks(513, 10.0, 8000.0, 128, 16000, norm="slaney", mel_scale="slaney")
    assert torch.equal(SP.timbre_derived_buffers()["mel.fb"], fb)
    # torch.stft centres a short window in the frame: spectrum of one frame with the padded window == torch.stft's
    x = torch.randn(1024)
    want = torch.stft(x, 1024, hop_length=320, win_length=640, window=torch.hann_window(640), center=False, return_complex=True)[:, 0]
    got = torch.fft.rfft(x * SP.centred_hann())
    assert float((got - want).abs().max()) < 1e-4


def test_full_size_15s_host_orchestration_vs_reference(emu, gold):
    """BASELINE config 5's full prompt size (15 s: 1498 fbank frames, 749 TDNN rows in 8 CAM segments, 751 mel frames)
    against the unmodified reference (tests/golden/speaker_full_15s.npz)."""
    from oracle import speaker as S
    g = gold("speaker_full_15s")
    wave = torch.cat([synth.synth_audio_16k(int(s), float(g["seconds"])) for s in g["seeds"]]).contiguous()
    assert np.abs(_style(emu, wave).numpy() - g["style"][0]).max() < 1e-4
    out, idx, z = _timbre(emu, wave)
    _, _, bounded = S.fsq4_quantize(z)
    safe = ((bounded - bounded.floor() - 0.5).abs() > 1e-3).all(dim=-1).numpy()
    assert safe.mean() > 0.9
    assert np.array_equal(idx.numpy()[safe], g["indices"][0, 0][safe])
    assert np.abs(out.numpy() - g["timbre"][0])[safe].max() < 1e-4

"""Generates tests/golden/style_vec.npz and tests/golden/timbre_latent.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_style          # build container only (needs /root/reference and torchaudio)

`InferenceWrapper.calculate_style_vec` (evaluations/infer_arvc.py:179-211) with the reference's own CAMPPlus module
(configs/hydra_arcs/sv/campplus.yaml) holding the seeded synthetic checkpoint, on seeded synthetic 16 kHz audio: one
5 s row, and a ragged batch of two rows.  Also records the fbank features of the first row (torchaudio's kaldi port).
`InferenceWrapper.calculate_timbre_latent` (:213-223) with the reference's own SpeakerEncoder
(configs/hydra_arcs/sv/sparktts_speaker_encoder.yaml, mel_fn = torchaudio MelSpectrogram) on the same audio."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness  # noqa: E402
from streamvoiceanon_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
WEIGHT_SEED = 1234


def main():
    ref_harness._paths()
    import torchaudio.compliance.kaldi as kaldi
    from evaluations.infer_arvc import InferenceWrapper
    from modules.campplus.DTDNN import CAMPPlus

    torch.set_num_threads(8)
    enc = CAMPPlus(feat_dim=80, embedding_size=192)
    enc.load_state_dict(synth.make_campplus_state_dict(WEIGHT_SEED), strict=True)
    enc.eval()
    w = object.__new__(InferenceWrapper)
    w.style_encoder = enc
    a = synth.synth_audio_16k(5100, 5.0)[None]
    b = synth.synth_audio_16k(5101, 3.0)[None]
    out = {"weight_seed": WEIGHT_SEED, "seed_a": 5100, "seed_b": 5101, "sec_a": 5.0, "sec_b": 3.0}
    with torch.no_grad():
        out["fbank_a"] = kaldi.fbank(a, num_mel_bins=80, dither=0, sample_frequency=16000).numpy()
        out["style_a"] = w.calculate_style_vec(a, torch.LongTensor([a.shape[1]])).numpy()
        batch = torch.zeros(2, a.shape[1])
        batch[0], batch[1, : b.shape[1]] = a[0], b[0]
        lens = torch.LongTensor([a.shape[1], b.shape[1]])
        out["style_batch"] = w.calculate_style_vec(batch, lens).numpy()
        out["batch_lens"] = lens.numpy()
    np.savez_compressed(GOLD / "style_vec.npz", **out)
    print("wrote", GOLD / "style_vec.npz", {k: getattr(v, "shape", v) for k, v in out.items()})

    import torchaudio
    from modules.bicodec_speaker_encoder.speaker_encoder import SpeakerEncoder
    mel_fn = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=640, hop_length=320, f_min=10.0,
                                                  f_max=None, n_mels=128, power=1.0, norm="slaney", mel_scale="slaney")
    tim = SpeakerEncoder(mel_fn=mel_fn, input_dim=128, out_dim=1024, latent_dim=128, token_num=32, fsq_levels=[4] * 6,
                         fsq_num_quantizers=1)
    missing, unexpected = tim.load_state_dict(synth.make_timbre_encoder_state_dict(WEIGHT_SEED), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("mel_transformer.", "project.", "speaker_encoder.bn.", "speaker_encoder.linear.",
                             "speaker_encoder.pool.")) for k in missing), missing        # not on the tokenize_wav path
    tim.eval()
    w.timbre_encoder = tim
    out = {"weight_seed": WEIGHT_SEED, "seed_a": 5100, "seed_b": 5101, "sec_a": 5.0, "sec_b": 3.0}
    with torch.no_grad():
        out["mel_a"] = mel_fn(a).squeeze(1).mT.numpy()
        out["timbre_a"] = w.calculate_timbre_latent(a, torch.LongTensor([a.shape[1]])).numpy()
        out["indices_a"] = tim.tokenize_wav(a, torch.LongTensor([a.shape[1]]))[1].numpy()
        out["timbre_batch"] = w.calculate_timbre_latent(batch, lens).numpy()
        out["indices_batch"] = tim.tokenize_wav(batch, lens)[1].numpy()
        out["batch_lens"] = lens.numpy()
    np.savez_compressed(GOLD / "timbre_latent.npz", **out)
    print("wrote", GOLD / "timbre_latent.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()

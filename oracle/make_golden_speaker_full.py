"""Generates tests/golden/speaker_full_15s.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_speaker_full        # build container only (needs /root/reference and torchaudio)

The speaker encoders at BASELINE config 5's full size: three 5 s references concatenated (15 s at 16 kHz = 1498 fbank
frames -> 749 TDNN rows in 8 CAM segments, 751 mel frames) through the reference's own `calculate_style_vec` and
`calculate_timbre_latent` / `tokenize_wav` (evaluations/infer_arvc.py:179-223) with its own CAMPPlus and SpeakerEncoder."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.make_golden_prompt import WEIGHT_SEED, build_wrapper  # noqa: E402
from streamvoiceanon_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
SEEDS, SECONDS = (5500, 5501, 5502), 5.0


def main():
    w, _, _ = build_wrapper()
    wave = torch.cat([synth.synth_audio_16k(s, SECONDS) for s in SEEDS])[None]
    lens = torch.LongTensor([wave.shape[1]])
    with torch.no_grad():
        style = w.calculate_style_vec(wave, lens)
        timbre = w.calculate_timbre_latent(wave, lens)
        indices = w.timbre_encoder.tokenize_wav(wave, lens)[1]
    np.savez_compressed(GOLD / "speaker_full_15s.npz", weight_seed=WEIGHT_SEED, seeds=np.array(SEEDS), seconds=SECONDS,
                        style=style.numpy(), timbre=timbre.numpy(), indices=indices.numpy())
    print("wrote speaker_full_15s.npz", tuple(style.shape), tuple(timbre.shape), tuple(indices.shape), wave.shape)


if __name__ == "__main__":
    main()

"""Generates tests/golden/vocoder_encode.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_vocenc          # build container only (needs /root/reference)

`FireflyArchitecture.encode` (modules/vqgan/modules/firefly.py:561-574) of the reference's own vocoder object -- the
prompt's "reference wave -> codec ids" step (`wav2target_fn`, evaluations/infer_arvc.py:168-171) -- on seeded synthetic
audio with the seeded synthetic checkpoint (decode-path dict merged with synth.make_vocoder_encoder_state_dict).  The FSQ
index arithmetic runs through the reference's vendored twin of vector-quantize-pytorch (oracle/refshim)."""
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
    torch.manual_seed(0)
    torch.set_num_threads(8)
    voc_sd = dict(synth.make_vocoder_state_dict(WEIGHT_SEED))
    voc_sd.update(synth.make_vocoder_encoder_state_dict(WEIGHT_SEED))
    _, _, voc, _ = ref_harness.build(synth.make_ar_state_dict(WEIGHT_SEED), synth.make_tokenizer_state_dict(WEIGHT_SEED),
                                     voc_sd, lambda step, slot, V: synth.noise_tape(7000, step)[slot])
    out = {}
    with torch.no_grad():
        for name, seed, frames in (("a", 1400, 24), ("b", 1401, 57)):
            wav = synth.synth_audio_44k(seed, 3.0)[: frames * 2048][None]
            (codes, quantized), lens = voc.encode(wav, torch.LongTensor([wav.shape[1]]))
            assert tuple(codes.shape) == (1, 8, frames) and int(lens[0]) == frames
            out[f"seed_{name}"], out[f"frames_{name}"], out[f"codes_{name}"] = seed, frames, codes.numpy().astype(np.int32)
    np.savez_compressed(GOLD / "vocoder_encode.npz", weight_seed=WEIGHT_SEED, **out)
    print("wrote", GOLD / "vocoder_encode.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

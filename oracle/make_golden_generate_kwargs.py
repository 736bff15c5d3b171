"""Generates tests/golden/ar_generate_kwargs.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_generate_kwargs        # build container only (needs /root/reference)

`ARVCWrapper.generate(..., temperature=0.9, top_p=0.85)` (modules/arvc_wrapper.py:82-98 -> dual_ar_stream.py:698-762) on
the inputs of tests/golden/ar_stream.npz: the reference samples the FIRST frame with the default sampling arguments
(:723 passes no kwargs) and every later frame with the caller's -- a quirk the engine preserves."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden import TAPE_SEED, WEIGHT_SEED, noise_fn  # noqa: E402
from streamvoiceanon_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
TEMPERATURE, TOP_P = 0.9, 0.85


def main():
    torch.set_num_threads(8)
    s = np.load(GOLD / "ar_stream.npz")
    model, _, _, tape = ref_harness.build(synth.make_ar_state_dict(WEIGHT_SEED), synth.make_tokenizer_state_dict(WEIGHT_SEED),
                                          synth.make_vocoder_state_dict(WEIGHT_SEED), noise_fn)
    style, timbre = synth.synth_speaker(int(s["spk_seed"]))
    with torch.no_grad():
        model.set_delay(delay=2)
        tape.step = -1
        out = model.generate(ref_content_codes=torch.from_numpy(s["ref_content"]), ref_audio_codes=torch.from_numpy(s["ref_audio"]),
                             src_content_codes=torch.from_numpy(s["src_content"])[:, :12], style_vectors=style,
                             timbre_latents=timbre, temperature=TEMPERATURE, top_p=TOP_P)
    np.savez_compressed(GOLD / "ar_generate_kwargs.npz", codes=out.numpy(), temperature=np.float32(TEMPERATURE),
                        top_p=np.float32(TOP_P), n_src=12, tape_seed=TAPE_SEED)
    print("wrote ar_generate_kwargs.npz", out.shape, out[0, :, :3].tolist())


if __name__ == "__main__":
    main()

"""Generates tests/golden/noise_mix.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_noise_mix          # build container only (needs /root/reference)

Calls `InferenceWrapper.apply_noise_mixing` (evaluations/infer_arvc.py:228-232; the method does not touch `self`) on
seeded inputs of the two shapes the reference mixes -- style_vectors [1,192] and timbre_latents [1,32,128]
(:419-421) -- and records the standard-normal draws it took from torch's global generator by replaying the seed."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def main():
    ref_harness._paths()
    from evaluations.infer_arvc import InferenceWrapper

    out = {}
    cases = (("style", (1, 192), 0.7, 11), ("timbre", (1, 32, 128), 0.7, 12), ("style_a0", (1, 192), 0.0, 13),
             ("timbre_a1", (1, 32, 128), 1.0, 14), ("ragged", (3, 37), 0.35, 15))
    for name, shape, alpha, seed in cases:
        g = torch.Generator().manual_seed(900 + seed)
        x = torch.randn(shape, generator=g) * 0.8 + 0.3
        torch.manual_seed(seed)
        y = InferenceWrapper.apply_noise_mixing(None, x, alpha)
        torch.manual_seed(seed)
        noise = torch.randn_like(x)                                # the draws the method just took
        out[f"x_{name}"], out[f"noise_{name}"], out[f"y_{name}"] = x.numpy(), noise.numpy(), y.numpy()
        out[f"alpha_{name}"], out[f"seed_{name}"] = np.float32(alpha), seed
    np.savez_compressed(GOLD / "noise_mix.npz", names=np.array([c[0] for c in cases]), **out)
    print("wrote", GOLD / "noise_mix.npz")


if __name__ == "__main__":
    main()

"""Generates tests/golden/config12_named.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_named           # build container only (needs /root/reference); ~3 minutes on 8 cores

BASELINE configs 1 and 2 on the inputs BASELINE.json names -- the reference's own test vectors
`test_waves/trump_0.wav` (source, 7.79 s stereo, 343 483 samples -> 167 frames) and `test_waves/azuma_0.wav` (reference
speaker, 7.13 s -> 153 prompt frames), copies of which sit beside the fixture as tests/golden/{trump_0,azuma_0}.wav
(input vectors held by the reference, not code):

  config 1  `InferenceWrapper.infer("test_waves/trump_0.wav", "test_waves/azuma_0.wav", delay=2, save_result=False)`
            (evaluations/infer_arvc.py:261-380), no compile, CPU
  config 2  `InferenceWrapper.stream_infer(same pair, decode_chunk_frames=1, delay=2, save_result=False)` with the CLI
            defaults (encode window 128, decode window 64, max prompt 256, max seq 768, buffer 32; :598-676, :706-711):
            168 chunks

through the reference's own top-level entry points with all five of its own modules (seeded synthetic checkpoints,
alpha = 1: no anonymisation noise), noise tape 7000.  Waveforms are stored as float32."""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden_prompt import WEIGHT_SEED, build_wrapper  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def main():
    w, tape, _ = build_wrapper()
    if not torch.cuda.is_available():                               # patch (ii) of SURVEY 8c-3, as in ref_harness
        torch.cuda.Event = ref_harness._Event
        torch.cuda.synchronize = lambda *a, **k: None
    w.compiled_speech_tokenizer_encode = w.speech_tokenizer.encode
    src, ref = str(GOLD / "trump_0.wav"), str(GOLD / "azuma_0.wav")
    with torch.no_grad():
        t0 = time.time()
        tape.step = -1
        wave = w.infer(src, ref, delay=2, save_result=False)
        t1 = time.time()
        tape.step = -1
        stream_wave = w.stream_infer(src, ref, decode_chunk_frames=1, delay=2, save_result=False)
        t2 = time.time()
    out = dict(weight_seed=WEIGHT_SEED, tape_seed=7000, wave=np.asarray(wave, dtype=np.float32),
               stream_wave=np.asarray(stream_wave, dtype=np.float32), stream_src_content=w.src_content_codes.numpy(),
               stream_pred_codes=w.pred_codes.numpy(), ref_content=w.ref_content_codes.numpy(),
               ref_audio=w.ref_audio_codes.numpy().astype(np.int32), style=w.style_vectors.numpy(), timbre=w.timbre_latents.numpy(),
               infer_seconds=t1 - t0, stream_seconds=t2 - t1)
    np.savez_compressed(GOLD / "config12_named.npz", **out)
    print("wrote", GOLD / "config12_named.npz", {k: getattr(v, "shape", v) for k, v in out.items()},
          "rms", float(np.sqrt((out["wave"] ** 2).mean())), float(np.sqrt((out["stream_wave"] ** 2).mean())))


if __name__ == "__main__":
    main()

"""Generates tests/golden/infer_config1.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_infer           # build container only (needs /root/reference and torchaudio)

BASELINE configs 1 and 2 as a user runs them -- from .wav FILES, through the reference's own top-level entry points with
the reference's own five modules (seeded synthetic checkpoints):

  * `InferenceWrapper.infer(src_path, [ref_a, ref_b], delay=2, alpha=0.7, save_result=False)`
    (evaluations/infer_arvc.py:261-380): offline conversion, two references, anonymisation mix -- once with the default
    "concat_mel" collation and once with "avg" (embeddings of each reference averaged, :282-307; unlike
    `calculate_prompt`, `infer` implements that branch completely);
  * `InferenceWrapper.stream_infer(src_path, ref_a, ..., decode_chunk_frames=1, delay=2, save_result=False)`
    (:598-676): the streaming loop including its file loading and left padding to whole chunks (small windows so that
    the CPU run stays short and the re-prompt path fires).

The .wav files are float32 PCM written from the seeded synthetic signals (the librosa shim reads them back bit-exactly).
The two `randn_like` draws of `infer`'s noise mix are recorded by replaying the seed."""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
from scipy.io import wavfile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden_prompt import WEIGHT_SEED, build_wrapper  # noqa: E402
from streamvoiceanon_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
MIX_SEED, ALPHA = 78, 0.7
SRC_SEED, SRC_SECONDS = 1010, 1.3
REF_SEEDS, REF_SECONDS = (5300, 5301), 1.1
STREAM_CFG = dict(encode_window_frames=24, decode_window_frames=16, max_prompt_frames=32, max_seq_frames=60, buffer_frames=6,
                  decode_chunk_frames=1, delay=2)


def main():
    w, tape, _ = build_wrapper()
    if not torch.cuda.is_available():                               # patch (ii) of SURVEY 8c-3, as in ref_harness
        torch.cuda.Event = ref_harness._Event
        torch.cuda.synchronize = lambda *a, **k: None
    w.compiled_speech_tokenizer_encode = w.speech_tokenizer.encode
    tmp = Path(tempfile.mkdtemp())
    src = synth.synth_audio_44k(SRC_SEED, SRC_SECONDS)
    wavfile.write(tmp / "src.wav", 44100, src.numpy())
    ref_paths = []
    for s in REF_SEEDS:
        wavfile.write(tmp / f"ref{s}.wav", 44100, synth.synth_audio_44k(s, REF_SECONDS).numpy())
        ref_paths.append(str(tmp / f"ref{s}.wav"))
    with torch.no_grad():
        tape.step = -1
        torch.manual_seed(MIX_SEED)
        wave = w.infer(str(tmp / "src.wav"), ref_paths, delay=2, alpha=ALPHA, save_result=False)
        torch.manual_seed(MIX_SEED)
        noise_style = torch.randn(1, 192)                            # the two draws infer just took, in order (:345-346)
        noise_timbre = torch.randn(1, 32, 128)
        tape.step = -1
        torch.manual_seed(MIX_SEED)                                  # same draws: the shapes are the same
        wave_avg = w.infer(str(tmp / "src.wav"), ref_paths, delay=2, alpha=ALPHA, spk_emb_collate_type="avg", save_result=False)
        tape.step = -1
        stream_wave = w.stream_infer(str(tmp / "src.wav"), ref_paths[0], save_result=False, alpha=1.0, **STREAM_CFG)
    out = dict(weight_seed=WEIGHT_SEED, mix_seed=MIX_SEED, alpha=np.float32(ALPHA), tape_seed=7000, src_seed=SRC_SEED,
               src_seconds=SRC_SECONDS, ref_seeds=np.array(REF_SEEDS), ref_seconds=REF_SECONDS, noise_style=noise_style.numpy(),
               noise_timbre=noise_timbre.numpy(), wave=np.asarray(wave, dtype=np.float32),
               wave_avg=np.asarray(wave_avg, dtype=np.float32),
               stream_wave=np.asarray(stream_wave, dtype=np.float32), stream_src_content=w.src_content_codes.numpy(),
               stream_pred_codes=w.pred_codes.numpy(), **{f"stream_{k}": np.array(v) for k, v in STREAM_CFG.items()})
    np.savez_compressed(GOLD / "infer_config1.npz", **out)
    print("wrote", GOLD / "infer_config1.npz", {k: getattr(v, "shape", v) for k, v in out.items()},
          "rms", float(np.sqrt((out["wave"] ** 2).mean())), float(np.sqrt((out["stream_wave"] ** 2).mean())))


if __name__ == "__main__":
    main()

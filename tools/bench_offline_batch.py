"""BASELINE config 3 on the GPU box: batched offline conversion of N synthetic 10 s utterances (each with its own 5 s
reference) through `InferenceWrapper.infer_batch`, beside N sequential `infer` calls (what the batch-1 reference does).
Prints one JSON line: wall seconds of both, audio seconds produced per wall second.

    python tools/bench_offline_batch.py [N=64] [sequential_sample=4]"""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from streamvoiceanon_b200 import InferenceWrapper, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_seq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    seed = 1234
    iw = InferenceWrapper.from_state_dicts(
        synth.make_ar_state_dict(seed), synth.make_tokenizer_state_dict(seed),
        {**synth.make_vocoder_state_dict(seed), **synth.make_vocoder_encoder_state_dict(seed)},
        synth.make_campplus_state_dict(seed), synth.make_timbre_encoder_state_dict(seed))
    srcs = [synth.synth_audio_44k(1000 + k, 10.0) for k in range(n)]
    refs = [synth.synth_audio_44k(5000 + k, 5.0) for k in range(n)]
    iw.infer_batch(srcs[:2], refs[:2], delay=2)                     # warm-up: workspaces, lazy set-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    waves = iw.infer_batch(srcs, refs, delay=2)
    t_batch = time.perf_counter() - t0
    audio_s = sum(w.shape[0] for w in waves) / 44100
    iw.infer(srcs[0], refs[0], delay=2)
    t0 = time.perf_counter()
    for k in range(n_seq):
        iw.infer(srcs[k], refs[k], delay=2)
    t_seq = (time.perf_counter() - t0) / n_seq
    print(json.dumps({"utterances": n, "audio_seconds": round(audio_s, 1), "infer_batch_wall_s": round(t_batch, 3),
                      "audio_seconds_per_wall_second": round(audio_s / t_batch, 1),
                      "sequential_infer_wall_s_per_utterance": round(t_seq, 3),
                      "sequential_estimate_wall_s": round(t_seq * n, 2), "speedup_vs_sequential": round(t_seq * n / t_batch, 2)}))


if __name__ == "__main__":
    main()

"""End-to-end example on the GPU box: the engine's `InferenceWrapper` built from (synthetic) reference-keyed state dicts,
`stream_infer` of a synthetic 10 s source against three 5 s references with the anonymisation mix (BASELINE config 5's
shape at chunk 2, or any chunk size), then the offline `infer` of the same pair.  Prints one JSON line with wall times.

    python tools/demo_stream_infer.py [decode_chunk_frames] [alpha]

With real checkpoints: pass `torch.load(...)` of the five files named in
configs/config_firefly_arvcasr_8192_delay0_8.yaml:43-57 to `InferenceWrapper.from_state_dicts` instead."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from streamvoiceanon_b200 import InferenceWrapper, synth  # noqa: E402


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
    seed = 1234
    t0 = time.perf_counter()
    iw = InferenceWrapper.from_state_dicts(
        synth.make_ar_state_dict(seed), synth.make_tokenizer_state_dict(seed),
        {**synth.make_vocoder_state_dict(seed), **synth.make_vocoder_encoder_state_dict(seed)},
        synth.make_campplus_state_dict(seed), synth.make_timbre_encoder_state_dict(seed))
    t_load = time.perf_counter() - t0
    src = synth.synth_audio_44k(1000, 10.0)
    refs = [synth.synth_audio_44k(5000 + i, 5.0) for i in range(3)]
    torch.manual_seed(0)
    out = {"decode_chunk_frames": chunk, "alpha": alpha, "load_s": round(t_load, 2)}
    for name in ("first", "second"):                     # the first call pays workspace growth and lazy set-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        wave = iw.stream_infer(src, refs, decode_chunk_frames=chunk, delay=2, alpha=alpha)
        dt = time.perf_counter() - t0
        out[f"stream_infer_{name}"] = {"seconds_out": round(wave.shape[0] / 44100, 2), "wall_s": round(dt, 3),
                                       "rtf_incl_prompt": round(dt / (wave.shape[0] / 44100), 4),
                                       "rms": round(float((wave ** 2).mean() ** 0.5), 4)}
    t0 = time.perf_counter()
    wave = iw.infer(src, refs, delay=2, alpha=alpha)
    out["infer"] = {"seconds_out": round(wave.shape[0] / 44100, 2), "wall_s": round(time.perf_counter() - t0, 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

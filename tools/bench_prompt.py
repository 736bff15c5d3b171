"""Setup-path timing on the GPU box: `PromptBuilder.calculate_prompt` (resample -> style vector -> timbre latents ->
noise mix -> codec ids -> content ids) for reference audio of a few lengths, per step, CUDA-event timed after warm-up,
plus the library's kernel-launch count per call.

    python tools/bench_prompt.py [seconds ...]          (default 5 10 15)

Not part of bench.py: the prompt runs once per stream (and the speaker encoders never again), so it is reported
beside the headline, not inside it."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from streamvoiceanon_b200 import ContentTokenizer, Vocoder, _lib, synth  # noqa: E402
from streamvoiceanon_b200.audio import Resampler  # noqa: E402
from streamvoiceanon_b200.prompt import PromptBuilder, apply_noise_mixing  # noqa: E402
from streamvoiceanon_b200.speaker import CAMPPlus, SpeakerEncoder, calculate_style_vec, calculate_timbre_latent  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.kernel_launches()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (_lib.kernel_launches() - n0) // reps


def main():
    secs = [float(x) for x in sys.argv[1:]] or [5.0, 10.0, 15.0]
    seed = 1234
    tok, voc, style, timbre = ContentTokenizer(), Vocoder(), CAMPPlus(), SpeakerEncoder()
    tok.load_state_dict(synth.make_tokenizer_state_dict(seed), strict=False)
    voc.load_state_dict({**synth.make_vocoder_state_dict(seed), **synth.make_vocoder_encoder_state_dict(seed)}, strict=False)
    style.load_state_dict(synth.make_campplus_state_dict(seed))
    timbre.load_state_dict(synth.make_timbre_encoder_state_dict(seed))
    pb = PromptBuilder(tok, voc, style, timbre)
    rs = Resampler(44100, 16000)
    for s in secs:
        ref = synth.synth_audio_44k(5000, s)[None].cuda()
        ref16 = rs(ref)
        lens16 = torch.LongTensor([ref16.shape[-1]])
        lens = torch.LongTensor([ref.shape[-1]])
        row = {"seconds": s}
        for name, fn in (("resample", lambda: rs(ref)),
                         ("style_vector", lambda: calculate_style_vec(style, ref16, lens16)),
                         ("timbre_latent", lambda: calculate_timbre_latent(timbre, ref16, lens16)),
                         ("noise_mix", lambda: apply_noise_mixing(torch.zeros(1, 32, 128, device="cuda"), 0.7)),
                         ("codec_ids", lambda: voc.encode(ref, lens)),
                         ("content_ids", lambda: tok.encode(ref, lens)),
                         ("calculate_prompt", lambda: pb.calculate_prompt(ref, 0.7))):
            ms, launches = timed(fn)
            row[name] = {"ms": round(ms, 3), "launches": int(launches)}
        print(json.dumps(row))


if __name__ == "__main__":
    main()

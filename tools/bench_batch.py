#!/usr/bin/env python
"""Concurrent-streams sweep of the lock-step batch path: ms per chunk step, RTF and frames/s for B streams on one GPU
(BASELINE metric "concurrent streams/GPU at RTF<1").  python tools/bench_batch.py [B ...]"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
FRAME_S = 2048 / 44100.0


def run(B, steps=12, warm=4, enc_win=128, dec_win=64, chunk=1, prompt_s=5.0, profile_steps=0, advance=0):
    from streamvoiceanon_b200 import BatchSession, ContentTokenizer, StreamSession, synth
    tok = ContentTokenizer()
    sessions = []
    ref_wave = synth.synth_audio_44k(5000, prompt_s)
    ref_wave = ref_wave[: (ref_wave.numel() // 2048) * 2048][None]
    n_ref = ref_wave.shape[1] // 2048
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    style, timbre = synth.synth_speaker(5000)
    for b in range(B):
        g = torch.Generator().manual_seed(99 + b)
        ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=g).int()
        s = StreamSession()
        s.set_sampling(0.7, 0.7, seed=7000 + b)
        s.set_prompt(ref_content[0], ref_audio.cuda(), style.cuda(), timbre.cuda(), 256, 2)
        sessions.append(s)
    batch = BatchSession(sessions)
    batch.setup(enc_win, dec_win, 768, 32, chunk)
    src = torch.stack([synth.synth_audio_44k(1000 + (b % 8), 2.0)[: 40 * 2048] for b in range(B)]).cuda()
    out = torch.empty(B, chunk * 2048, device="cuda")
    it = 0

    def step():
        nonlocal it
        i = it % (40 // chunk)
        batch.process_chunk(src[:, i * chunk * 2048:(i + 1) * chunk * 2048], out)
        it += 1
    for _ in range(warm + advance):       # `advance`: grow the KV caches before measuring (S_valid += 2 per step)
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    ms = e0.elapsed_time(e1) / steps
    batch.set_timing(True)
    st = []
    for _ in range(5):
        step()
        st.append(batch.last_timing())
    med = [sorted(x[j] for x in st)[2] for j in range(3)]
    if profile_steps:                     # ncu --profile-from-start off: only these steps are profiled
        batch.set_timing(False)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(profile_steps):
            step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    batch.close()
    for s in sessions:
        s.close()
    return dict(streams=B, chunk=chunk, ms_per_step=ms, wall_ms_per_step=wall, rtf=ms / 1e3 / (chunk * FRAME_S),
                frames_per_s=B * chunk / (ms / 1e3), stage_ms={"E": med[0], "A": med[1], "V": med[2]})


def main():
    from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, Vocoder, synth
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
    ContentTokenizer().load_state_dict(synth.make_tokenizer_state_dict(1234), strict=False)
    Vocoder().load_state_dict(synth.make_vocoder_state_dict(1234), strict=False)
    Bs = [int(x) for x in sys.argv[1:]] or [1, 4, 8, 16, 32]
    for B in Bs:
        print(json.dumps(run(B)), flush=True)


if __name__ == "__main__":
    main()

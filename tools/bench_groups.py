#!/usr/bin/env python
"""Experiment: G groups of B/G streams, each group a lock-step batch on its own engine instance and CUDA stream, stepped
back to back from one host thread so that the latency-bound decode launches of one group overlap the GEMMs of another.
python tools/bench_groups.py B G [B G ...]"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
FRAME_S = 2048 / 44100.0


def make_group(B, seed0, enc_win=128, dec_win=64, chunk=1, prompt_s=5.0):
    from streamvoiceanon_b200 import ARVCWrapper, BatchSession, ContentTokenizer, StreamSession, Vocoder, synth
    from streamvoiceanon_b200 import engine as E
    E._ENGINES[torch.cuda.current_device()] = E.Engine(torch.cuda.current_device())      # a fresh engine for this group
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
    tok = ContentTokenizer()
    tok.load_state_dict(synth.make_tokenizer_state_dict(1234), strict=False)
    voc = Vocoder()
    voc.load_state_dict(synth.make_vocoder_state_dict(1234), strict=False)
    ref_wave = synth.synth_audio_44k(5000, prompt_s)
    ref_wave = ref_wave[: (ref_wave.numel() // 2048) * 2048][None]
    n_ref = ref_wave.shape[1] // 2048
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    style, timbre = synth.synth_speaker(5000)
    sessions = []
    for b in range(B):
        g = torch.Generator().manual_seed(99 + seed0 + b)
        ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=g).int()
        s = StreamSession()
        s.set_sampling(0.7, 0.7, seed=7000 + seed0 + b)
        s.set_prompt(ref_content[0], ref_audio.cuda(), style.cuda(), timbre.cuda(), 256, 2)
        sessions.append(s)
    batch = BatchSession(sessions)
    batch.setup(enc_win, dec_win, 768, 32, chunk)
    src = torch.stack([synth.synth_audio_44k(1000 + (b % 8), 2.0)[: 40 * 2048] for b in range(B)]).cuda()
    out = torch.empty(B, chunk * 2048, device="cuda")
    return dict(batch=batch, src=src, out=out, keep=(ar, tok, voc, sessions), stream=torch.cuda.Stream())


def run(B, G, steps=12, warm=4):
    groups = [make_group(B // G, 1000 * g) for g in range(G)]
    torch.cuda.synchronize()
    it = 0

    def step():
        nonlocal it
        i = it % 40
        for gr in groups:
            with torch.cuda.stream(gr["stream"]):
                gr["batch"].process_chunk(gr["src"][:, i * 2048:(i + 1) * 2048], gr["out"])
        it += 1

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    for gr in groups:
        gr["batch"].close()
    return dict(streams=B, groups=G, ms_per_step=ms, rtf=ms / 1e3 / FRAME_S, frames_per_s=B / (ms / 1e3))


def main():
    a = [int(x) for x in sys.argv[1:]] or [128, 1, 128, 2]
    for B, G in zip(a[::2], a[1::2]):
        print(json.dumps(run(B, G)), flush=True)


if __name__ == "__main__":
    main()

"""Perf mode (svanon_set_precision 1: fp16 single-pass tensor-core GEMMs, the reference's own GPU precision) against parity
mode (fp32-grade 3xTF32), SURVEY section 8c contract: "in bf16/fp16 perf mode report id-agreement rate and teacher-forced
max-abs logit error instead of claiming bit-exactness".  Prints one JSON line:

  content_id_agreement      stage E: BSQ ids of `utts` synthetic utterances of `seconds` s, perf vs parity (the ids depend on
                            the audio only, so every frame is compared)
  codec_id_agreement_tf     stage A, many-stream decode: B streams decode `steps` frames in both modes from identical prompts
                            and sampler seeds; every frame is compared UNTIL a stream's first disagreement, i.e. only frames
                            whose whole history is identical in both runs (= teacher forcing by construction); the rate is
                            agreeing codebook entries / compared entries; also the mean frame of first divergence
  fast_logits_max_abs_err   max |logit_perf - logit_parity| over the 1000-way fast-head logits of the first decoded frame of
                            stream 0 (identical history; codebooks after the first disagreeing token are left out)
  vocoder_snr_db            stage V: waveform of the same codes in both modes

    python tools/eval_perf_mode.py [B=32] [steps=40]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from streamvoiceanon_b200 import ARVCWrapper, BatchSession, ContentTokenizer, StreamSession, Vocoder, _lib, synth  # noqa: E402


def decode_run(tok, B, steps, precision, eng):
    eng.set_precision(precision)
    ref_wave = synth.synth_audio_44k(5000, 5.0)
    ref_wave = ref_wave[: (ref_wave.numel() // 2048) * 2048][None]
    n_ref = ref_wave.shape[1] // 2048
    eng.set_precision(0)                                   # identical prompts: the prompt ids come from parity mode
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    eng.set_precision(precision)
    style, timbre = synth.synth_speaker(5000)
    sessions = []
    for b in range(B):
        g = torch.Generator().manual_seed(99 + b)
        s = StreamSession()
        s.set_sampling(0.7, 0.7, seed=7000 + b)
        s.set_prompt(ref_content[0], torch.randint(0, 1000, (1, 8, n_ref), generator=g).int().cuda(), style.cuda(), timbre.cuda(), 256, 2)
        sessions.append(s)
    batch = BatchSession(sessions)
    batch.set_encoder_mode(3)                              # stateful encoder: cheap, and deterministic given the audio
    batch.setup(128, 64, 768, 32, 1)
    batch.set_ar_path(1)
    src = torch.stack([synth.synth_audio_44k(1000 + (b % 8), 4.0)[: (steps + 2) * 2048] for b in range(B)]).cuda()
    lib = _lib.load()
    _lib.check(lib.svanon_ar_debug_logits(eng.handle, 1))
    logits = None
    out = torch.empty(B, 2048, device="cuda")
    for i in range(steps + 2):
        # stage E runs in the run's own precision too: frames are compared only while the content ids of both runs agree
        batch.process_chunk(src[:, i * 2048:(i + 1) * 2048], out)
        if i == 2:
            fl = torch.empty(8, 1000)
            _lib.check(lib.svanon_ar_read_debug(eng.handle, None, None, fl.data_ptr()))
            logits = fl.clone()
    _lib.check(lib.svanon_ar_debug_logits(eng.handle, 0))
    hist = [s.history() for s in sessions]
    batch.close()
    for s in sessions:
        s.close()
    eng.set_precision(0)
    return [h[0] for h in hist], [h[1] for h in hist], logits


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
    tok = ContentTokenizer()
    tok.load_state_dict(synth.make_tokenizer_state_dict(1234), strict=False)
    voc = Vocoder()
    voc.load_state_dict(synth.make_vocoder_state_dict(1234), strict=False)
    eng = ar._engine
    res = {"streams": B, "steps": steps}
    # ---- stage E
    utts, seconds = 16, 6.0
    wavs = torch.stack([synth.synth_audio_44k(2000 + i, seconds + 0.1)[: int(seconds * 44100) // 2048 * 2048] for i in range(utts)]).cuda()
    lens = torch.LongTensor([wavs.shape[1]] * utts).cuda()
    eng.set_precision(0)
    ids0, _ = tok.encode(wavs, lens)
    eng.set_precision(1)
    ids1, _ = tok.encode(wavs, lens)
    eng.set_precision(0)
    bits = (ids0 ^ ids1)
    nbits = sum(int(((bits >> k) & 1).sum()) for k in range(13))
    res["content_id_agreement"] = float((ids0 == ids1).float().mean())
    res["content_bit_agreement"] = 1.0 - nbits / (13.0 * ids0.numel())
    res["content_frames_compared"] = int(ids0.numel())
    # ---- stage A (decoders fed by the stateful encoder; both runs' E is whatever the precision gives -- compare only while equal)
    src0, pred0, log0 = decode_run(tok, B, steps, 0, eng)
    src1, pred1, log1 = decode_run(tok, B, steps, 1, eng)
    agree = total = 0
    first = []
    for b in range(B):
        n = min(pred0[b].shape[1], pred1[b].shape[1])
        same_src = (src0[b][: n + 2] == src1[b][: n + 2])
        div = n
        for t in range(n):
            if not bool(same_src[: t + 3].all()):          # content ids already differ: history no longer identical
                div = t
                break
            eq = (pred0[b][:, t] == pred1[b][:, t])
            agree += int(eq.sum())
            total += 8
            if not bool(eq.all()):
                div = t
                break
        first.append(div)
    res["codec_id_agreement_tf"] = agree / max(total, 1)
    res["codec_entries_compared"] = total
    res["mean_first_divergence_frame"] = sum(first) / len(first)
    res["streams_never_diverged"] = sum(1 for f in first if f >= steps)
    ok = 8
    tok0, tok1 = pred0[0][:, 0], pred1[0][:, 0]
    for k in range(8):
        if int(tok0[k]) != int(tok1[k]):
            ok = k + 1
            break
    res["fast_logits_max_abs_err"] = float((log0[:ok] - log1[:ok]).abs().max())
    res["fast_logits_codebooks_compared"] = ok
    res["fast_logits_abs_max"] = float(log0.abs().max())
    # ---- stage V
    g = torch.Generator().manual_seed(4242)
    codes = torch.randint(0, 1000, (1, 8, 64), generator=g).cuda()
    eng.set_precision(0)
    w0 = voc.decode_codes(codes)
    eng.set_precision(1)
    w1 = voc.decode_codes(codes)
    eng.set_precision(0)
    noise = float(((w0 - w1) ** 2).mean())
    res["vocoder_snr_db"] = float(10 * torch.log10((w0 ** 2).mean() / max(noise, 1e-30)))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

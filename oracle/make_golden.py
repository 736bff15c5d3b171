"""Generates tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden            # ~2-3 minutes on 8 cores

Every fixture holds the seeds of its inputs (weights seed, audio seeds, tape seed) and
the reference's outputs; inputs are regenerated from `streamvoiceanon_b200.synth` by
the tests, so the fixtures stay small.  The reference's own modules produce every
output stored here (through oracle/ref_harness.py); the oracle is only compared, never
used to produce a fixture.
"""
from __future__ import annotations

import hashlib
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness  # noqa: E402
from streamvoiceanon_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
WEIGHT_SEED = 1234
TAPE_SEED = 7000


def noise_fn(step, slot, V):
    return synth.noise_tape(TAPE_SEED, step)[slot]


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes()[:4096])
    return h.hexdigest()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    GOLD.mkdir(parents=True, exist_ok=True)
    t0 = time.time()
    ar_sd = synth.make_ar_state_dict(WEIGHT_SEED)
    tok_sd = synth.make_tokenizer_state_dict(WEIGHT_SEED)
    voc_sd = synth.make_vocoder_state_dict(WEIGHT_SEED)
    print(f"weights {time.time() - t0:.1f}s")
    model, tok, voc, tape = ref_harness.build(ar_sd, tok_sd, voc_sd, noise_fn)
    digests = dict(ar=sd_digest(ar_sd), tok=sd_digest(tok_sd), voc=sd_digest(voc_sd))
    np.savez(GOLD / "weights_digest.npz", **{k: np.array(v) for k, v in digests.items()}, seed=WEIGHT_SEED)

    with torch.no_grad():
        # ---------------------------------------------------------------- E
        wav = synth.synth_audio_44k(1000, 2.0)[: 40 * 2048][None]
        ids, flen = tok.encode(wav, torch.LongTensor([wav.shape[1]]))
        mel = tok.spec_transform(wav)
        feat = tok.backbone(mel)
        z = tok.quantizer.pre_module(tok.quantizer.downsample(feat))
        proj = tok.quantizer.residual_bsq.rvqs[0].project_in(z.mT)
        print("E ids", ids.shape, ids[0, 0, :8].tolist(), "min|proj|", proj.abs().min().item())
        np.savez_compressed(GOLD / "encoder_40f.npz", audio_seed=1000, n_samples=wav.shape[1],
                            ids=ids.numpy(), mel_tail=mel[0, :, -8:].numpy(), feat_tail=feat[0, :, -8:].numpy(),
                            z_tail=z[0, :, -4:].numpy(), proj=proj[0].numpy())
        # streaming-size window: 128 frames, first 100 frames silent (zero ring at stream start)
        win = torch.zeros(1, 128 * 2048)
        win[:, -28 * 2048:] = synth.synth_audio_44k(1001, 2.0)[: 28 * 2048]
        ids_w, _ = tok.encode(win, torch.LongTensor([win.shape[1]]))
        np.savez_compressed(GOLD / "encoder_window128.npz", audio_seed=1001, live_frames=28, ids=ids_w.numpy())

        # ---------------------------------------------------------------- V
        g = torch.Generator().manual_seed(4242)
        codes = torch.randint(0, 1000, (1, 8, 20), generator=g)
        zq = voc.quantizer.decode(codes)
        wave = voc.head(zq)
        print("V wave", wave.shape, "rms", wave.pow(2).mean().sqrt().item(), "absmax", wave.abs().max().item())
        np.savez_compressed(GOLD / "vocoder_20f.npz", codes_seed=4242, codes=codes.numpy(),
                            z_tail=zq[0, :, -8:].numpy(), wave=wave[0, 0].numpy())

        # ---------------------------------------------------------------- A (streaming state machine)
        g = torch.Generator().manual_seed(31337)
        T = 24
        ref_content = torch.randint(0, 8192, (1, T), generator=g)
        ref_audio = torch.randint(0, 1000, (1, 8, T), generator=g).int()
        src_content = torch.randint(0, 8192, (1, 16), generator=g)
        style, timbre = synth.synth_speaker(5000)
        model.set_delay(delay=2)
        tape.step = -1
        model.prefill_prompt(ref_content, ref_audio, style, timbre)
        model.prefill_src_condition4delay(src_content[:, :2])
        codes_out, poss, logits1, hidden = [], [], [], []
        import modules.dual_ar_stream as das
        for t in range(2, 16):
            c, pos = model.decode_one(src_content[:, t: t + 1])
            codes_out.append(c.clone().numpy())
            poss.append(int(pos))
        print("A codes", np.stack(codes_out)[:3, :, 0].tolist(), "pos", poss[:3])
        np.savez_compressed(GOLD / "ar_stream.npz", seed=31337, T=T, delay=2, tape_seed=TAPE_SEED, spk_seed=5000,
                            ref_content=ref_content.numpy(), ref_audio=ref_audio.numpy(), src_content=src_content.numpy(),
                            codes=np.stack(codes_out), pos=np.array(poss))

        # teacher-forced logits: one decode_one_token_ar call with captured fast logits
        captured = {}
        real_fast = model.decoder.model.forward_generate_fast

        def cap_fast(x, input_pos=None):
            out = real_fast(x, input_pos)
            captured.setdefault("fast", []).append(out[0, -1].clone())
            return out

        real_fg = model.decoder.model.forward_generate

        def cap_fg(x, input_pos=None, kv_pos=None, vq_masks=None):
            out = real_fg(x, input_pos, kv_pos, vq_masks)
            captured["slow_logits"] = out.logits[0, -1].clone()
            captured["hidden"] = out.hidden_states[0, -1].clone()
            return out

        model.decoder.model.forward_generate_fast = cap_fast
        model.decoder.model.forward_generate = cap_fg
        tape.step = -1
        model.prefill_prompt(ref_content, ref_audio, style, timbre)
        pre_hidden = captured["hidden"].numpy().copy()
        pre_logits = captured["slow_logits"].numpy().copy()
        captured.clear()
        model.prefill_src_condition4delay(src_content[:, :2])
        captured.clear()
        c, pos = model.decode_one(src_content[:, 2:3])
        np.savez_compressed(GOLD / "ar_logits.npz", prefill_hidden=pre_hidden, prefill_logits=pre_logits,
                            hidden=captured["hidden"].numpy(), slow_logits=captured["slow_logits"].numpy(),
                            fast_logits=torch.stack(captured["fast"]).numpy(), codes=c.numpy())
        model.decoder.model.forward_generate_fast = real_fast
        model.decoder.model.forward_generate = real_fg

        # ---------------------------------------------------------------- A (offline generate)
        tape.step = -1
        out = model.generate(ref_content_codes=ref_content, ref_audio_codes=ref_audio, src_content_codes=src_content[:, :10],
                             style_vectors=style, timbre_latents=timbre)
        print("A generate", out.shape)
        np.savez_compressed(GOLD / "ar_generate.npz", codes=out.numpy())

        # ---------------------------------------------------------------- loop (unmodified process_one_chunk)
        # NOTE: the reference crashes unless the prompt holds >= decode_window_frames-1 frames
        # (infer_arvc.py:567-583 pads the window with at most len(ref_audio_codes) frames).
        for name, kw, n_chunks, n_ref in (
            ("stream_default", dict(encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32), 26, 72),
            ("stream_reprompt", dict(encode_window_frames=32, decode_window_frames=16, max_seq_frames=56, buffer_frames=8), 22, 24),
        ):
            t1 = time.time()
            style, timbre = synth.synth_speaker(5001)
            ref_wave = synth.synth_audio_44k(5001, 3.5)[: n_ref * 2048][None]
            g = torch.Generator().manual_seed(99)
            ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=g).int()
            w = ref_harness.make_inference_wrapper(model, tok, voc, style, timbre, ref_audio)
            tape.step = -1
            w.prefill_prompt([ref_wave], max_prompt_frames=256, delay=2)
            w.setup_stream_caches(decode_chunk_frames=1, **kw)
            src = synth.synth_audio_44k(1002, 1.5)[: n_chunks * 2048].view(n_chunks, 2048)
            waves = []
            for i in range(n_chunks):
                waves.append(w.process_one_chunk(src[i][None]).clone())
            waves = torch.cat(waves, dim=-1)
            print(name, "chunks", n_chunks, f"{time.time() - t1:.1f}s", "codes", tuple(w.pred_codes.shape),
                  "wave rms", waves.pow(2).mean().sqrt().item())
            np.savez_compressed(GOLD / f"{name}.npz", ref_seed=5001, src_seed=1002, codes_seed=99, n_chunks=n_chunks,
                                tape_seed=TAPE_SEED, delay=2, n_ref=n_ref, **{k: np.array(v) for k, v in kw.items()},
                                ref_content=w.ref_content_codes.numpy(), src_content=w.src_content_codes.numpy(),
                                pred_codes=w.pred_codes.numpy(), wave=waves[0].numpy())
    print(f"done {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()

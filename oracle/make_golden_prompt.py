"""Generates tests/golden/prompt_config5.npz and stream_config5.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_prompt          # build container only (needs /root/reference and torchaudio)

`InferenceWrapper.calculate_prompt` (evaluations/infer_arvc.py:382-441) with three seeded synthetic references,
alpha = 0.7, "concat_mel" -- the prompt of BASELINE config 5 (anonymisation) -- on an InferenceWrapper whose five models
are the reference's own modules holding the seeded synthetic checkpoints.  The two `randn_like` draws of the noise mix
are recorded by replaying the seed."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness  # noqa: E402
from streamvoiceanon_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
WEIGHT_SEED, MIX_SEED, ALPHA = 1234, 77, 0.7
REF_SEEDS, REF_SECONDS = (5200, 5201, 5202), 1.6


def build_wrapper():
    """The reference's `InferenceWrapper` around the reference's own five modules holding the synthetic checkpoints (no
    method replaced).  Returns (wrapper, noise tape, (model, tokenizer, vocoder))."""
    torch.set_num_threads(8)
    voc_sd = dict(synth.make_vocoder_state_dict(WEIGHT_SEED))
    voc_sd.update(synth.make_vocoder_encoder_state_dict(WEIGHT_SEED))
    model, tok, voc, tape = ref_harness.build(synth.make_ar_state_dict(WEIGHT_SEED), synth.make_tokenizer_state_dict(WEIGHT_SEED),
                                           voc_sd, lambda step, slot, V: synth.noise_tape(7000, step)[slot])
    import torchaudio
    from evaluations.infer_arvc import InferenceWrapper
    from modules.bicodec_speaker_encoder.speaker_encoder import SpeakerEncoder
    from modules.campplus.DTDNN import CAMPPlus

    style_enc = CAMPPlus(feat_dim=80, embedding_size=192)
    style_enc.load_state_dict(synth.make_campplus_state_dict(WEIGHT_SEED), strict=True)
    mel_fn = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=640, hop_length=320, f_min=10.0,
                                                  f_max=None, n_mels=128, power=1.0, norm="slaney", mel_scale="slaney")
    timbre_enc = SpeakerEncoder(mel_fn=mel_fn, input_dim=128, out_dim=1024, latent_dim=128, token_num=32,
                                fsq_levels=[4] * 6, fsq_num_quantizers=1)
    missing, unexpected = timbre_enc.load_state_dict(synth.make_timbre_encoder_state_dict(WEIGHT_SEED), strict=False)
    assert not unexpected
    w = object.__new__(InferenceWrapper)
    w.device = torch.device("cpu")
    w.sr = 44100
    w.model, w.speech_tokenizer, w.firefly = model, tok, voc
    w.style_encoder, w.timbre_encoder = style_enc.eval(), timbre_enc.eval()
    return w, tape, (model, tok, voc)


def main():
    w, tape, (model, tok, voc) = build_wrapper()
    refs = [synth.synth_audio_44k(s, REF_SECONDS)[None] for s in REF_SEEDS]
    with torch.no_grad():
        torch.manual_seed(MIX_SEED)
        codes, content, style, timbre, ref = w.calculate_prompt(refs, alpha=ALPHA, spk_emb_collate_type="concat_mel")
        torch.manual_seed(MIX_SEED)
        noise_style = torch.randn_like(style)                       # the two draws calculate_prompt just took, in order
        noise_timbre = torch.randn_like(timbre)
    out = dict(weight_seed=WEIGHT_SEED, mix_seed=MIX_SEED, alpha=np.float32(ALPHA), ref_seeds=np.array(REF_SEEDS),
               ref_seconds=REF_SECONDS, ref_audio_codes=codes.numpy().astype(np.int32), ref_content_codes=content.numpy(),
               style_vectors=style.numpy(), timbre_latents=timbre.numpy(), noise_style=noise_style.numpy(),
               noise_timbre=noise_timbre.numpy(), n_samples=ref.shape[-1])
    np.savez_compressed(GOLD / "prompt_config5.npz", **out)
    print("wrote", GOLD / "prompt_config5.npz", {k: getattr(v, "shape", v) for k, v in out.items()})

    # ---- the same prompt through the UNMODIFIED streaming loop, chunk = 2 (BASELINE config 5): prefill_prompt (which
    # calls calculate_prompt and truncates to max_prompt_frames, infer_arvc.py:462-489), setup_stream_caches,
    # process_one_chunk; small windows so that the re-prompt path (:547-564) fires with two-frame chunks
    ws = ref_harness.make_inference_wrapper(model, tok, voc, None, None, None)       # installs the CPU Event stubs
    del ws.calculate_prompt                                                          # back to the class's own method
    ws.style_encoder, ws.timbre_encoder = w.style_encoder, w.timbre_encoder
    cfg = dict(encode_window_frames=32, decode_window_frames=16, max_seq_frames=72, buffer_frames=8, decode_chunk_frames=2)
    n_chunks, max_prompt, delay = 12, 48, 2
    tape.step = -1
    with torch.no_grad():
        torch.manual_seed(MIX_SEED)
        ws.prefill_prompt(refs, max_prompt_frames=max_prompt, delay=delay, alpha=ALPHA)
        ws.setup_stream_caches(**cfg)
        src = synth.synth_audio_44k(1003, 1.5)[: n_chunks * 4096].view(n_chunks, 4096)
        waves = torch.cat([ws.process_one_chunk(src[i][None]).clone() for i in range(n_chunks)], dim=-1)
    out = dict(weight_seed=WEIGHT_SEED, mix_seed=MIX_SEED, alpha=np.float32(ALPHA), tape_seed=7000, src_seed=1003,
               n_chunks=n_chunks, max_prompt_frames=max_prompt, delay=delay, **{k: np.array(v) for k, v in cfg.items()},
               src_content=ws.src_content_codes.numpy(), pred_codes=ws.pred_codes.numpy(), wave=waves[0].numpy())
    np.savez_compressed(GOLD / "stream_config5.npz", **out)
    print("wrote", GOLD / "stream_config5.npz", {k: getattr(v, "shape", v) for k, v in out.items()},
          "wave rms", float(waves.pow(2).mean().sqrt()))


if __name__ == "__main__":
    main()

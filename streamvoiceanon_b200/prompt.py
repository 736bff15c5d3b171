"""Prompt-side helpers of the setup path that sit between the speaker encoders and
`prefill_prompt`: the anonymisation noise mix of `InferenceWrapper.apply_noise_mixing`
(evaluations/infer_arvc.py:228-232, call sites :419-421)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import Engine, ptr, _cuda_stream_ptr


@torch.no_grad()
def apply_noise_mixing(tensor: torch.Tensor, alpha: float, noise: torch.Tensor | None = None, device=None) -> torch.Tensor:
    """`alpha * tensor + (1 - alpha) * (randn_like(tensor) * tensor.std() + tensor.mean())` on the GPU.

    Same argument meaning as the reference method.  `noise` defaults to `torch.randn_like(tensor)` drawn from torch's
    global generator exactly where the reference draws it, so a shared `torch.manual_seed` gives the same mix; pass the
    draws explicitly to replay a tape.  Returns a tensor of the input's shape on the input's device (a host input is
    staged through the C ABI)."""
    eng = Engine.get(device if device is not None else (tensor.device if tensor.is_cuda else None))
    x = tensor.detach().to(torch.float32).contiguous()
    if noise is None:
        noise = torch.randn_like(tensor)
    if tuple(noise.shape) != tuple(tensor.shape):
        raise ValueError(f"noise shape {tuple(noise.shape)} != tensor shape {tuple(tensor.shape)}")
    nz = noise.detach().to(torch.float32).contiguous()
    out = torch.empty_like(x)
    _lib.check(eng.lib.svanon_noise_mix(eng.handle, ptr(x), ptr(nz), x.numel(), C.c_float(float(alpha)), ptr(out),
                                        C.c_void_p(_cuda_stream_ptr())))
    return out.to(tensor.dtype)


class PromptBuilder:
    """`InferenceWrapper.calculate_prompt` (evaluations/infer_arvc.py:382-441) over the engine: reference waves at
    44.1 kHz -> (ref_audio_codes [1,8,T] i32, ref_content_codes [1,T] i64, style_vectors [1,192], timbre_latents
    [1,32,128], ref_wav [1,n]) -- the five tensors `prefill_prompt` consumes (:462-489).  Every step is a library call:
    svanon_resample (44.1 -> 16 kHz), svanon_style_vector, svanon_timbre_latent, svanon_noise_mix (style first, then
    timbre: the order of the reference's two `randn_like` draws, :419-421), svanon_voc_encode, svanon_enc_encode.

    `calculate_prompt` offers the "concat_mel" collation only, like the reference in effect: its "avg" branch falls through
    to code that reads a variable the branch never sets (:411-417) and cannot run with more than one reference.  The
    reference's offline `infer` carries a complete copy of that branch (:282-307: embeddings of each reference on its own,
    averaged; ids still from the concatenation) -- `allow_avg=True` is that copy."""

    def __init__(self, speech_tokenizer, firefly, style_encoder, timbre_encoder, sr: int = 44100, resample_freq: int = 16000):
        self.speech_tokenizer, self.firefly = speech_tokenizer, firefly
        self.style_encoder, self.timbre_encoder = style_encoder, timbre_encoder
        self._rates, self._resampler = (sr, resample_freq), None

    def _resample(self, wave):
        if self._resampler is None:                      # built on first use: needs the engine (a GPU)
            from .audio import Resampler
            self._resampler = Resampler(*self._rates)
        return self._resampler(wave)

    @torch.no_grad()
    def calculate_prompt(self, ref_wav_tensors, alpha: float = 1.0, spk_emb_collate_type: str = "concat_mel",
                         noise_style: torch.Tensor | None = None, noise_timbre: torch.Tensor | None = None,
                         allow_avg: bool = False):
        from .speaker import calculate_style_vec, calculate_timbre_latent
        refs = list(ref_wav_tensors) if isinstance(ref_wav_tensors, (list, tuple)) else [ref_wav_tensors]
        avg = spk_emb_collate_type == "avg" and len(refs) > 1
        if avg and not allow_avg:
            raise NotImplementedError('spk_emb_collate_type="avg" cannot run in the reference with more than one '
                                      "reference wave (evaluations/infer_arvc.py:411-417); use \"concat_mel\"")
        dev = torch.device("cuda", self.style_encoder._engine.device)
        ref = (torch.cat(refs, dim=-1) if len(refs) > 1 else refs[0]).to(dev, torch.float32)
        if avg:
            styles, timbres = [], []
            for r in refs:
                r16 = self._resample(r.to(dev, torch.float32))
                l16 = torch.LongTensor([r16.shape[-1]])
                styles.append(calculate_style_vec(self.style_encoder, r16, l16))
                timbres.append(calculate_timbre_latent(self.timbre_encoder, r16, l16))
            style = torch.mean(torch.stack(styles, dim=0), dim=0)
            timbre = torch.mean(torch.stack(timbres, dim=0), dim=0)
        else:
            ref16 = self._resample(ref)
            lens16 = torch.LongTensor([ref16.shape[-1]])
            style = calculate_style_vec(self.style_encoder, ref16, lens16)
            timbre = calculate_timbre_latent(self.timbre_encoder, ref16, lens16)
        style = apply_noise_mixing(style, alpha, noise_style)              # draws (if any) in the reference's order
        timbre = apply_noise_mixing(timbre, alpha, noise_timbre)
        lens = torch.LongTensor([ref.shape[-1]])
        (codes, _), _ = self.firefly.encode(ref, lens)
        content, _ = self.speech_tokenizer.encode(ref, lens)
        return codes, content.squeeze(0), style, timbre, ref

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

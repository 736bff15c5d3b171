"""`torch.library` custom ops over the C ABI, so that the reference's own `torch.compile` calls keep working.

The reference compiles two of the three model surfaces in CALLER code we do not edit:
`torch.compile(self.firefly.head, fullgraph=True, mode="reduce-overhead")` (evaluations/infer_arvc.py:128-134) and the
same for `self.speech_tokenizer.encode` (:136-142); the GUI always does (real-time-gui.py:54-57).  With
`fullgraph=True` a plain ctypes call is a graph break, i.e. an error.  Registered as custom ops with fake (meta)
implementations, the engine entries are opaque graph nodes for Dynamo, and because the C ABI is stream-ordered, does
not allocate after warm-up and never synchronises for device buffers, the launches are CUDA-graph capturable, which is
what `reduce-overhead` does with them.  Device tensors only (the compiled callers run on the GPU)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import Engine, ptr, _cuda_stream_ptr


def _eng(t: torch.Tensor) -> Engine:
    if not t.is_cuda:
        raise RuntimeError("svanon_b200 custom ops take CUDA tensors")
    return Engine.get(t.device)


@torch.library.custom_op("svanon_b200::enc_encode", mutates_args=())
def enc_encode(wave: torch.Tensor) -> torch.Tensor:
    """FireflyArchitecture.encode (firefly_encoder.py:553-566) for full-length rows: wave [B, L] f32 -> ids [1, B, L//2048]."""
    eng = _eng(wave)
    w = wave.float().contiguous()
    B, L = w.shape
    ids = torch.empty(1, B, L // 2048, dtype=torch.int64, device=w.device)
    if ids.numel():
        _lib.check(eng.lib.svanon_enc_encode_batch(eng.handle, ptr(w), B, L, ptr(ids), C.c_void_p(_cuda_stream_ptr())))
    return ids


@enc_encode.register_fake
def _(wave):
    return wave.new_empty((1, wave.shape[0], wave.shape[1] // 2048), dtype=torch.int64)


@torch.library.custom_op("svanon_b200::voc_head", mutates_args=())
def voc_head(z: torch.Tensor) -> torch.Tensor:
    """HiFiGANGenerator.forward (firefly.py:280-293): z [B, 512, L] -> wave [B, 1, 512 L]."""
    eng = _eng(z)
    B, Cc, L = z.shape
    zl = z.transpose(1, 2).float().contiguous()
    wave = torch.empty(B, 1, 512 * L, dtype=torch.float32, device=z.device)
    for b in range(B):
        _lib.check(eng.lib.svanon_voc_head(eng.handle, ptr(zl[b]), L, ptr(wave[b]), C.c_void_p(_cuda_stream_ptr())))
    return wave


@voc_head.register_fake
def _(z):
    return z.new_empty((z.shape[0], 1, 512 * z.shape[2]), dtype=torch.float32)


@torch.library.custom_op("svanon_b200::voc_quantizer_decode", mutates_args=())
def voc_quantizer_decode(codes: torch.Tensor) -> torch.Tensor:
    """DownsampleFiniteScalarQuantize.decode (fsq.py:112-116): codes [B, 8, T] -> z [B, 4T, 512] (channels-last)."""
    eng = _eng(codes)
    B, G, T = codes.shape
    c = codes.to(torch.int64).contiguous()
    z = torch.empty(B, 4 * T, 512, dtype=torch.float32, device=codes.device)
    for b in range(B):
        _lib.check(eng.lib.svanon_voc_quantizer_decode(eng.handle, ptr(c[b]), T, ptr(z[b]), C.c_void_p(_cuda_stream_ptr())))
    return z


@voc_quantizer_decode.register_fake
def _(codes):
    return codes.new_empty((codes.shape[0], 4 * codes.shape[2], 512), dtype=torch.float32)


@torch.library.custom_op("svanon_b200::voc_decode", mutates_args=())
def voc_decode(codes: torch.Tensor) -> torch.Tensor:
    """code2wav_fn (evaluations/infer_arvc.py:173-176): codes [B, 8, T] -> wave [B, 1, 2048 T]."""
    eng = _eng(codes)
    B, G, T = codes.shape
    c = codes.to(torch.int64).contiguous()
    wave = torch.empty(B, 1, 2048 * T, dtype=torch.float32, device=codes.device)
    for b in range(B):
        _lib.check(eng.lib.svanon_voc_decode(eng.handle, ptr(c[b]), T, ptr(wave[b]), C.c_void_p(_cuda_stream_ptr())))
    return wave


@voc_decode.register_fake
def _(codes):
    return codes.new_empty((codes.shape[0], 1, 2048 * codes.shape[2]), dtype=torch.float32)

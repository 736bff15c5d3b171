"""Process-wide engine handle (one per GPU) and weight upload helpers."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import torch

from . import _lib

_ENGINES: Dict[int, "Engine"] = {}


def _cuda_stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("svanon_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = device
        h = C.c_void_p()
        _lib.check(self.lib.svanon_engine_create(device, C.byref(h)))
        self.handle = h
        self.loaded = {0: False, 1: False, 2: False, 3: False, 4: False}
        self._digest = {}
        import os
        mode = os.environ.get("SVANON_GEMM_MODE")        # 1 = fp32 CUDA cores, 2 = tcgen05 3xTF32 (library default)
        if mode is not None:
            _lib.check(self.lib.svanon_set_gemm_mode(int(mode)))
        prec = os.environ.get("SVANON_PRECISION")      # 0 parity (3xTF32, default), 1 perf (fp16 single-pass tensor-core GEMMs)
        if prec is not None:
            _lib.check(self.lib.svanon_set_precision(int(prec)))
        pdl = os.environ.get("SVANON_PDL")
        if pdl is not None:
            _lib.check(self.lib.svanon_set_pdl(int(pdl)))

    def set_precision(self, mode: int):
        """0: parity mode (fp32-grade 3xTF32 GEMMs, ids bit-exact); 1: perf mode (fp16 single-pass tensor-core GEMMs, the
        reference's own GPU precision).  Process-wide (include/svanon.h)."""
        _lib.check(self.lib.svanon_set_precision(int(mode)))

    @staticmethod
    def get(device=None) -> "Engine":
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else 0
        if device not in _ENGINES:
            _ENGINES[device] = Engine(device)
        return _ENGINES[device]

    def load_tensor(self, model: int, name: str, t: torch.Tensor):
        if self.loaded[model]:              # same checkpoint loaded again (load_state_dict verified the digest): nothing to do
            return
        t = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        _lib.check(self.lib.svanon_load_tensor(self.handle, model, name.encode(), ptr(t), t.dim(), shape))

    def load_state_dict(self, model: int, sd: Dict[str, torch.Tensor], wanted) -> Tuple[list, list]:
        """Uploads the tensors `wanted(key)` accepts; returns (missing-from-wanted-set is checked at
        finalize by the library, unexpected = keys nobody wanted)."""
        unexpected, take = [], []
        for k, v in sd.items():
            if not torch.is_tensor(v) or not (v.is_floating_point()):
                continue
            if v.dim() == 0 or v.dim() > 4:
                continue
            if wanted(k):
                take.append((k, v))
            else:
                unexpected.append(k)
        # One engine per GPU holds ONE checkpoint per model (streams keep pointers into it).  Constructing the model
        # objects a second time with the same checkpoint (a second InferenceWrapper, a GUI reload) is a no-op; a
        # different checkpoint needs a new process.
        import hashlib
        h = hashlib.sha1()
        for k, v in sorted(take, key=lambda kv: kv[0]):
            f = v.detach().reshape(-1)
            h.update(f"{k}|{tuple(v.shape)}|{float(f.double().sum()):.9e}|{float(f[0]):.9e}|{float(f[-1]):.9e};".encode())
        digest = h.hexdigest()
        if self.loaded[model]:
            if digest != self._digest.get(model):
                raise RuntimeError("this engine already holds different weights for this model: one engine per GPU and "
                                   "process holds one checkpoint (streams point into it); start a new process to load another")
            return unexpected
        self._digest[model] = digest
        for k, v in take:
            self.load_tensor(model, k, v)
        return unexpected

    def finalize(self, model: int):
        if self.loaded[model]:
            return
        _lib.check(self.lib.svanon_finalize_weights(self.handle, model))
        self.loaded[model] = True


def rope_table(seq_len: int, n_elem: int = 64, base: float = 10000.0) -> torch.Tensor:
    """precompute_freqs_cis (modules/dual_ar_stream.py:993-1001, windowed_transformer.py:356-365),
    computed with the same torch CPU ops, rounded to bf16 and widened back to fp32."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: n_elem // 2].float() / n_elem))
    freqs = torch.outer(torch.arange(seq_len), freqs)
    cis = torch.polar(torch.ones_like(freqs), freqs)
    return torch.stack([cis.real, cis.imag], dim=-1).to(torch.bfloat16).float()


def slaney_fbanks(n_freqs=1025, f_min=0.0, f_max=22050.0, n_mels=160, sample_rate=44100) -> torch.Tensor:
    """The filterbank `LogMelSpectrogram.__init__` builds with torchaudio.functional.melscale_fbanks(
    norm="slaney", mel_scale="slaney") (modules/vqgan/spectrogram.py:93-106); [n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0

    def hz_to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    m_pts = torch.linspace(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2)
    f_pts = f_sp * m_pts
    log_t = m_pts >= min_log_mel
    f_pts[log_t] = min_log_hz * torch.exp(logstep * (m_pts[log_t] - min_log_mel))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2: n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)

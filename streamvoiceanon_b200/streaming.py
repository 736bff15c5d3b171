"""The per-chunk loop (`InferenceWrapper.process_one_chunk`, evaluations/infer_arvc.py:492-596) as ONE
library call per chunk: wave ring, E window, warm-up, A, re-prompt, V window, tail select all run
stream-ordered inside `svanon_stream_process_chunk` with no host synchronisation until the result is
copied out."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib
from .engine import Engine, ptr, _cuda_stream_ptr


def new_sampler_seed() -> int:
    """A fresh seed for the library's counter-based Exp(1) generator, drawn from torch's GLOBAL generator: the
    reference samples from that generator (modules/dual_ar_stream.py:1095), fresh per run and per stream, and
    `torch.manual_seed` makes it reproducible -- the same holds here."""
    return int(torch.randint(0, 2 ** 62, (), dtype=torch.int64))


class StreamSession:
    def __init__(self, device=None, max_seq_len: int = 2048):
        self._engine = Engine.get(device)
        for m in (0, 1, 2):
            if not self._engine.loaded[m]:
                raise RuntimeError("load the AR, tokenizer and vocoder weights first (ARVCWrapper / ContentTokenizer / "
                                   "Vocoder .load_state_dict)")
        h = C.c_void_p()
        _lib.check(self._engine.lib.svanon_stream_create(self._engine.handle, max_seq_len, C.byref(h)))
        self._h = h
        self.chunk = 1
        self._noise_fn: Optional[Callable] = None
        self._step = 0
        self.set_sampling(seed=new_sampler_seed())     # every new stream samples differently unless told otherwise

    def set_noise_fn(self, fn, step0: int = 0):
        self._noise_fn, self._step = fn, step0

    def set_sampling(self, temperature=0.7, top_p=0.7, seed=0):
        _lib.check(self._engine.lib.svanon_ar_set_sampling(self._h, temperature, top_p, seed))

    def set_prompt(self, ref_content_codes, ref_audio_codes, style_vectors, timbre_latents, max_prompt_frames=256, delay=2):
        """Tail of InferenceWrapper.prefill_prompt (infer_arvc.py:468-489)."""
        rc = ref_content_codes.reshape(-1).to(torch.int64).contiguous()
        ra = ref_audio_codes.reshape(8, -1).to(torch.int32).contiguous()
        sv = style_vectors.reshape(-1).float().contiguous()
        tl = timbre_latents.reshape(32, 128).float().contiguous()
        _lib.check(self._engine.lib.svanon_stream_set_prompt(self._h, ptr(rc), ptr(ra), rc.numel(), ptr(sv), ptr(tl),
                                                             max_prompt_frames, delay, C.c_void_p(_cuda_stream_ptr())))
        self.delay = delay
        self._step += 1
        self._n_src = 0
        self._prefilled = False

    def setup(self, encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32,
              decode_chunk_frames=1):
        """InferenceWrapper.setup_stream_caches (infer_arvc.py:443-460)."""
        _lib.check(self._engine.lib.svanon_stream_setup(self._h, encode_window_frames, decode_window_frames,
                                                        max_seq_frames, buffer_frames, decode_chunk_frames))
        self.chunk = decode_chunk_frames
        self.max_seq_frames = max_seq_frames
        self._n_src = 0
        self._prefilled = False

    def _noise(self):
        """Mirrors the oracle's step counter: one step per prefill / decode_one_token_ar call."""
        self._n_src += self.chunk
        if self._n_src < self.delay:
            return None
        if not self._prefilled and self.delay != 0:
            self._prefilled = True
            self._step += 1
            return None
        rows = []
        for _ in range(self.chunk):
            if self._noise_fn is not None:
                rows.append(torch.stack([self._noise_fn(self._step, s, 1000)[:1000] for s in range(1, 9)]))
            self._step += 1
        pos = int(self._engine.lib.svanon_ar_position(self._h)) + 2 * self.chunk - 1
        if pos // 2 >= self.max_seq_frames:
            self._step += 2 if self.delay > 0 else 1          # re-prompt: prefill_prompt + prefill_delay
        return torch.stack(rows).float().contiguous() if rows else None

    def process_chunk(self, wave_chunk: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """InferenceWrapper.process_one_chunk (infer_arvc.py:492-596).  `wave_chunk` [chunk*2048] on the host or the
        device; the result lands where `out` lives (default: same place as the input)."""
        w = wave_chunk.reshape(-1).float().contiguous()
        if out is None:
            out = torch.empty_like(w)
        noise = self._noise()
        _lib.check(self._engine.lib.svanon_stream_process_chunk(self._h, ptr(w), w.numel(),
                                                                ptr(noise) if noise is not None else None, ptr(out),
                                                                C.c_void_p(_cuda_stream_ptr())))
        return out

    def set_vocoder_mode(self, incremental: bool = True):
        """True (default): incremental vocoder; False: recompute the decode window every chunk like the reference."""
        _lib.check(self._engine.lib.svanon_stream_set_vocoder_mode(self._h, int(incremental)))

    def set_encoder_mode(self, incremental=True):
        """True / 1 (default): keep the conv-stack outputs of the window between chunks (ring-buffer state); False / 0:
        re-encode the whole window every chunk like the reference; 2: additionally continue the newest frames from
        per-layer conv history (what batches of >= 8 streams do on their own).  Same result.
        3 (before the first chunk): the STATEFUL encoder -- a different function: ids of the reference's offline `encode()`
        of the stream so far instead of its 128-frame window re-encode (include/svanon.h); 0.23 GFLOP per frame."""
        _lib.check(self._engine.lib.svanon_stream_set_encoder_mode(self._h, int(incremental)))

    def set_timing(self, enable: bool = True):
        _lib.check(self._engine.lib.svanon_stream_set_timing(self._h, int(enable)))

    def last_timing(self):
        """(E, A, V) device milliseconds of the last non-warm-up chunk."""
        ms = (C.c_float * 3)()
        _lib.check(self._engine.lib.svanon_stream_last_timing(self._h, ms))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def history(self, cap: int = 4096):
        src = torch.empty(cap, dtype=torch.int64)
        pred = torch.empty(8 * cap, dtype=torch.int64)
        ns, npred = C.c_int(0), C.c_int(0)
        _lib.check(self._engine.lib.svanon_stream_history(self._h, ptr(src), C.byref(ns), ptr(pred), C.byref(npred), cap))
        return src[: ns.value].clone(), pred[: 8 * npred.value].view(8, npred.value).clone()

    def close(self):
        if self._h is not None:
            self._engine.lib.svanon_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchSession:
    """N streams advanced in lock-step, one library call per chunk for all of them (BASELINE configs 3-4).

    The reference has no equivalent (it is batch-1: `max_batch_size=1`, evaluations/infer_arvc.py:56); the contract
    is that stream i of the batch produces exactly what `StreamSession` i produces alone.  Usage: create the
    StreamSessions, `set_prompt` each (same delay), then `BatchSession(sessions).setup(...)` and
    `process_chunk(waves [n, chunk*2048])` per chunk."""

    def __init__(self, sessions):
        if not sessions:
            raise ValueError("empty batch")
        self.sessions = list(sessions)
        self._engine = self.sessions[0]._engine
        arr = (C.c_void_p * len(self.sessions))(*[s._h for s in self.sessions])
        h = C.c_void_p()
        _lib.check(self._engine.lib.svanon_batch_create(self._engine.handle, arr, len(self.sessions), C.byref(h)))
        self._h = h
        self.chunk = 1

    def setup(self, encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32,
              decode_chunk_frames=1):
        _lib.check(self._engine.lib.svanon_batch_setup(self._h, encode_window_frames, decode_window_frames, max_seq_frames,
                                                       buffer_frames, decode_chunk_frames))
        self.chunk = decode_chunk_frames
        for s in self.sessions:
            s.chunk = decode_chunk_frames
            s.max_seq_frames = max_seq_frames
            s._n_src = 0
            s._prefilled = False

    @classmethod
    def merged(cls, a: "BatchSession", b: "BatchSession") -> "BatchSession":
        """One lock-step batch out of two that are both past their warm-up chunks (same settings): a's streams followed by
        b's, every stream with the state it had (svanon_batch_merge).  `a` and `b` are closed."""
        if a._engine is not b._engine:
            raise ValueError("batches of different engines")
        h = C.c_void_p()
        _lib.check(a._engine.lib.svanon_batch_merge(a._h, b._h, C.byref(h), C.c_void_p(_cuda_stream_ptr())))
        m = cls.__new__(cls)
        m.sessions = a.sessions + b.sessions
        m._engine, m._h, m.chunk = a._engine, h, a.chunk
        a.close()
        b.close()
        return m

    @classmethod
    def selected(cls, a: "BatchSession", keep) -> "BatchSession":
        """The members `keep` (indices into a.sessions, strictly increasing) of a batch that is past its warm-up chunks as a
        batch of their own, every stream with the state it had (svanon_batch_select): the batch a server continues with
        after streams have left.  `a` is closed; the sessions left out stay open (the caller owns them)."""
        keep = [int(k) for k in keep]
        arr = (C.c_int * len(keep))(*keep)
        h = C.c_void_p()
        _lib.check(a._engine.lib.svanon_batch_select(a._h, arr, len(keep), C.byref(h), C.c_void_p(_cuda_stream_ptr())))
        m = cls.__new__(cls)
        m.sessions = [a.sessions[k] for k in keep]
        m._engine, m._h, m.chunk = a._engine, h, a.chunk
        a.close()
        return m

    def set_ar_path(self, path: int):
        """0: persistent kernel for 1/2/4 streams, many-stream kernels otherwise; 1: always the many-stream kernels."""
        _lib.check(self._engine.lib.svanon_batch_set_ar_path(self._h, path))

    def process_chunk(self, wave_chunks: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = len(self.sessions)
        w = wave_chunks.reshape(n, -1).float().contiguous()
        if out is None:
            out = torch.empty_like(w)
        noises = [s._noise() for s in self.sessions]
        noise = None
        if any(z is not None for z in noises):
            if any(z is None for z in noises):
                raise RuntimeError("either every stream of a batch has a noise tape or none has")
            noise = torch.stack(noises).float().contiguous()
            if w.is_cuda:
                noise = noise.to(w.device)
        _lib.check(self._engine.lib.svanon_batch_process_chunk(self._h, ptr(w), w.shape[1],
                                                               ptr(noise) if noise is not None else None, ptr(out),
                                                               C.c_void_p(_cuda_stream_ptr())))
        return out

    def set_encoder_mode(self, incremental=True):
        _lib.check(self._engine.lib.svanon_batch_set_encoder_mode(self._h, int(incremental)))

    def set_timing(self, enable: bool = True):
        _lib.check(self._engine.lib.svanon_batch_set_timing(self._h, int(enable)))

    def last_timing(self):
        ms = (C.c_float * 3)()
        _lib.check(self._engine.lib.svanon_batch_last_timing(self._h, ms))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def close(self):
        if self._h is not None:
            self._engine.lib.svanon_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Host-side mirror of `modules.arvc_wrapper.ARVCWrapper` over the C ABI.

Same constructor signature, method names, argument meaning and error behaviour as the
reference class (modules/arvc_wrapper.py:7-126) so that `hydra.utils.instantiate` of
configs/hydra_arcs/vc/firefly_arvc_bsq_8192_delay0_8.yaml (with `_target_` pointed here, or
with this package shadowing `modules.arvc_wrapper`) and every call site in
evaluations/infer_arvc.py / real-time-gui.py keep working.  All arithmetic happens in
libsvanon_b200.so; this file only marshals tensors.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from typing import Callable, Optional

import torch

from . import _lib
from .engine import Engine, ptr, rope_table, _cuda_stream_ptr

_IncompatibleKeys = namedtuple("_IncompatibleKeys", ["missing_keys", "unexpected_keys"])


class DualARModelArgs:
    """Stand-in for modules.dual_ar_stream.DualARModelArgs (:100-129): holds the YAML fields;
    the engine is compiled for the shipped configuration and checks it."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.delay = kw.get("delay", 0)


class DualARTransformer:
    """Stand-in for modules.dual_ar_stream.DualARTransformer (:411-457): config carrier only."""

    def __init__(self, config):
        self.config = config


class DualARWrapper:
    """Stand-in for modules.dual_ar_stream.DualARWrapper (:605-637): exposes `delay` and
    `original_delay`, which evaluations/infer_arvc.py:476-479 reads."""

    def __init__(self, model):
        self.model = model
        d = getattr(model.config, "delay", 0)
        self.delay = d
        self.original_delay = d

    def set_delay(self, delay: int):
        delay = int(delay)
        if isinstance(self.original_delay, int):
            print("Setting delay is not supported to a model with a single delay value, ignoring operation...")
            return
        self.original_delay = self.model.config.delay
        self.delay = delay
        print(f"Setting delay to {self.delay} frames")


def _default_decoder():
    return DualARWrapper(DualARTransformer(DualARModelArgs(delay=list(range(9)))))


_WANTED_PREFIXES = ("embedding.weight", "decoder.model.codebook_embeddings.", "decoder.model.layers.",
                    "decoder.model.norm.", "decoder.model.output.", "decoder.model.fast_embeddings.",
                    "decoder.model.fast_layers.", "decoder.model.fast_norm.", "decoder.model.fast_output.",
                    "decoder.wait4start_embedding.", "decoder.wait4end_embedding.", "context_in.", "style_in.")
_IGNORED = ("decoder.model.embeddings.weight",)     # unused at inference


class ARVCWrapper:
    def __init__(self, embedding=None, decoder=None, context_dim: int = 128, style_dim: int = 192,
                 model_dim: int = 768, spk_condition: bool = True, device: Optional[int] = None):
        if (context_dim, style_dim, model_dim) != (128, 192, 768) or not spk_condition:
            raise ValueError("svanon_b200 is built for context_dim=128, style_dim=192, model_dim=768, spk_condition=True")
        self.embedding = embedding
        self.decoder = decoder if decoder is not None else _default_decoder()
        self.compiled_fn = None
        self.spk_condition = spk_condition
        self._engine = Engine.get(device)
        self._stream = None
        self._max_seq_len = 2048
        self._noise_fn: Optional[Callable] = None
        self._step = 0
        self.training = False

    # ---- nn.Module-ish no-ops the callers use (infer_arvc.py:60-65)
    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def half(self):
        return self

    def parameters(self):
        return iter(())

    def compile_ar_decode_fn(self):
        """modules/arvc_wrapper.py:25-41 compiles decode_one_token_ar with inductor; the engine's decode
        step already is a single persistent kernel, so there is nothing to compile."""
        self.compiled_fn = None

    # ---- weights
    def load_state_dict(self, sd, strict: bool = False):
        eng = self._engine
        unexpected = eng.load_state_dict(_lib.MODEL_AR, sd, lambda k: k.startswith(_WANTED_PREFIXES))
        unexpected = [k for k in unexpected if k not in _IGNORED]
        eng.load_tensor(_lib.MODEL_AR, "decoder.model.freqs_cis", rope_table(2048))
        eng.load_tensor(_lib.MODEL_AR, "decoder.model.fast_freqs_cis", rope_table(8))
        eng.finalize(_lib.MODEL_AR)     # raises RuntimeError naming the first missing tensor
        if strict and unexpected:
            raise RuntimeError(f"Unexpected key(s) in state_dict: {unexpected}")
        return _IncompatibleKeys([], unexpected)

    # ---- caches / delay
    def setup_caches(self, max_batch_size: int = 1, max_seq_len: int = 2048, dtype=None):
        """modules/arvc_wrapper.py:43-44 -> dual_ar_stream.py:225-243,459-475.  One KV slab per stream;
        `dtype` is accepted for signature compatibility (the parity build keeps K/V in fp32)."""
        if max_batch_size != 1:
            raise ValueError("ARVCWrapper is a single-stream surface (reference: max_batch_size=1, infer_arvc.py:56); "
                             "use streamvoiceanon_b200.BatchSession for concurrent streams")
        self._max_seq_len = int(max_seq_len)
        self._new_stream()

    def _new_stream(self):
        lib = self._engine.lib
        if self._stream is not None:
            lib.svanon_stream_destroy(self._stream)
        h = C.c_void_p()
        _lib.check(lib.svanon_stream_create(self._engine.handle, self._max_seq_len, C.byref(h)))
        self._stream = h
        # the reference samples from torch's global generator (dual_ar_stream.py:1095): a new stream draws the seed of
        # the library's counter-based generator from it (torch.manual_seed reproduces a run; no two streams are alike)
        from .streaming import new_sampler_seed
        _lib.check(lib.svanon_ar_set_sampling(h, 0.7, 0.7, new_sampler_seed()))
        d = self.decoder.delay
        if isinstance(d, int):
            _lib.check(lib.svanon_ar_set_delay(self._stream, d))

    def _need_stream(self):
        if self._stream is None:
            self._new_stream()
        return self._stream

    def set_delay(self, **kwargs):
        self.decoder.set_delay(**kwargs)
        if isinstance(self.decoder.delay, int):
            _lib.check(self._engine.lib.svanon_ar_set_delay(self._need_stream(), self.decoder.delay))

    # ---- sampling noise (tests): fn(step, slot, V) -> Exp(1) tensor, shared with the oracle
    def set_noise_fn(self, fn: Optional[Callable], step0: int = 0):
        self._noise_fn = fn
        self._step = step0

    def set_sampling(self, temperature: float = 0.7, top_p: float = 0.7, seed: int = 0):
        _lib.check(self._engine.lib.svanon_ar_set_sampling(self._need_stream(), temperature, top_p, seed))

    def _noise_for_step(self, step: int):
        if self._noise_fn is None:
            return None
        return torch.stack([self._noise_fn(step, s, 1000)[:1000] for s in range(1, 9)]).float().contiguous()

    # ---- reference surface
    def prefill_prompt(self, ref_content_codes, ref_audio_codes, style_vectors, timbre_latents):
        """modules/arvc_wrapper.py:100-112."""
        if ref_content_codes.shape[0] != 1:
            raise AssertionError("batch size must be 1")
        rc = ref_content_codes[0].to(torch.int64).contiguous()
        ra = ref_audio_codes[0].to(torch.int32).contiguous()
        sv = style_vectors.reshape(-1).float().contiguous()
        tl = timbre_latents.reshape(32, 128).float().contiguous()
        _lib.check(self._engine.lib.svanon_ar_prefill_prompt(self._need_stream(), ptr(rc), ptr(ra), rc.numel(),
                                                             ptr(sv), ptr(tl), C.c_void_p(_cuda_stream_ptr())))
        self._step += 1

    def prefill_src_condition4delay(self, src_content_codes):
        """modules/arvc_wrapper.py:114-119 (asserts src length == delay, dual_ar_stream.py:805)."""
        sc = src_content_codes.reshape(-1).to(torch.int64).contiguous()
        assert sc.numel() == self.decoder.delay
        _lib.check(self._engine.lib.svanon_ar_prefill_delay(self._need_stream(), ptr(sc), sc.numel(),
                                                            C.c_void_p(_cuda_stream_ptr())))
        self._step += 1

    def decode_one(self, src_content_codes):
        """modules/arvc_wrapper.py:121-126 -> (codes int32 [8,1], last position 0-d tensor)."""
        cid = src_content_codes.reshape(-1)[:1].to(torch.int64).contiguous()
        dev = cid.device if cid.is_cuda else torch.device("cuda", self._engine.device)
        out = torch.empty(8, 1, dtype=torch.int32, device=dev)
        noise = self._noise_for_step(self._step)
        pos = C.c_int32(0)
        _lib.check(self._engine.lib.svanon_ar_decode_one(self._need_stream(), ptr(cid),
                                                         ptr(noise) if noise is not None else None, ptr(out),
                                                         C.byref(pos), C.c_void_p(_cuda_stream_ptr())))
        self._step += 1
        # the position is host-side bookkeeping (no device read); the caller only does `current_pos // 2 >= max_seq_frames`
        # (infer_arvc.py:547), which a 0-d CPU tensor serves without a device round trip
        return out, torch.tensor(pos.value)

    def generate(self, ref_content_codes, ref_audio_codes, src_content_codes, style_vectors, timbre_latents,
                 **sampling_kwargs):
        """modules/arvc_wrapper.py:82-98 -> [1, 8, Ts] int32."""
        # per-call sampling arguments: the reference applies them from the SECOND frame on (its prefill call passes none,
        # dual_ar_stream.py:723); `repetition_penalty` never has an effect there (previous_tokens is always None)
        unknown = set(sampling_kwargs) - {"temperature", "top_p", "repetition_penalty"}
        if unknown:
            raise TypeError(f"generate() got unexpected sampling arguments {sorted(unknown)}")
        custom = "temperature" in sampling_kwargs or "top_p" in sampling_kwargs
        if custom:
            _lib.check(self._engine.lib.svanon_ar_set_generate_sampling(
                self._need_stream(), float(sampling_kwargs.get("temperature", 0.7)), float(sampling_kwargs.get("top_p", 0.7))))
        rc = ref_content_codes[0].to(torch.int64).contiguous()
        ra = ref_audio_codes[0].to(torch.int32).contiguous()
        sc = src_content_codes[0].to(torch.int64).contiguous()
        sv = style_vectors.reshape(-1).float().contiguous()
        tl = timbre_latents.reshape(32, 128).float().contiguous()
        Ts = sc.numel()
        dev = torch.device("cuda", self._engine.device)
        out = torch.empty(8, Ts, dtype=torch.int32, device=dev)
        noise = None
        if self._noise_fn is not None:
            noise = torch.stack([self._noise_for_step(self._step + i) for i in range(Ts)]).contiguous()
        try:
            _lib.check(self._engine.lib.svanon_ar_generate(self._need_stream(), ptr(rc), ptr(ra), rc.numel(), ptr(sc), Ts,
                                                           ptr(sv), ptr(tl), ptr(noise) if noise is not None else None,
                                                           ptr(out), C.c_void_p(_cuda_stream_ptr())))
        finally:
            if custom:
                _lib.check(self._engine.lib.svanon_ar_set_generate_sampling(self._need_stream(), -1.0, -1.0))
        self._step += Ts
        return out[None]

    # ---- test hooks
    def debug_logits(self, enable: bool = True):
        _lib.check(self._engine.lib.svanon_ar_debug_logits(self._engine.handle, int(enable)))

    def read_debug(self):
        slow = torch.empty(8192)
        hidden = torch.empty(768)
        fast = torch.empty(8, 1000)
        _lib.check(self._engine.lib.svanon_ar_read_debug(self._engine.handle, ptr(slow), ptr(hidden), ptr(fast)))
        return slow, hidden, fast

    def __del__(self):
        try:
            if self._stream is not None:
                self._engine.lib.svanon_stream_destroy(self._stream)
        except Exception:
            pass

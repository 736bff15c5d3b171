"""Host-side mirrors of the two `FireflyArchitecture` objects the hot path touches.

  ContentTokenizer  <-> modules.vqgan.modules.firefly_encoder.FireflyArchitecture  (only `encode`, :553-566)
  Vocoder           <-> modules.vqgan.modules.firefly.FireflyArchitecture          (`quantizer.decode` + `head`)

Constructor signatures follow the reference (`backbone, head, quantizer, spec_transform`, all ignored: the
engine is compiled for the shipped YAMLs) so hydra instantiation keeps working.  Arithmetic is in
libsvanon_b200.so.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple

import torch

from . import _lib
from .engine import Engine, ptr, rope_table, slaney_fbanks, _cuda_stream_ptr

_IncompatibleKeys = namedtuple("_IncompatibleKeys", ["missing_keys", "unexpected_keys"])


class _Shim:
    training = False

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def half(self):
        return self

    def parameters(self):
        return iter(())


class ContentTokenizer(_Shim):
    _WANTED = ("backbone.", "quantizer.downsample.", "quantizer.pre_module.layers.", "quantizer.pre_module.norm.",
               "quantizer.residual_bsq.rvqs.0.project_in.")

    def __init__(self, backbone=None, head=None, quantizer=None, spec_transform=None, device=None):
        self._engine = Engine.get(device)
        self.downsample_factor = 4

    def load_state_dict(self, sd, strict: bool = False):
        eng = self._engine
        unexpected = eng.load_state_dict(_lib.MODEL_TOKENIZER, sd, lambda k: k.startswith(self._WANTED))
        # persistent buffer of the checkpoint when present, else rebuilt like the reference
        fc = sd.get("quantizer.pre_module.freqs_cis")
        eng.load_tensor(_lib.MODEL_TOKENIZER, "quantizer.pre_module.freqs_cis",
                        fc.float() if fc is not None else rope_table(2048))
        eng.load_tensor(_lib.MODEL_TOKENIZER, "spec_transform.fb", slaney_fbanks())
        # the stateful encoder (EncoderStream / encoder mode 3) runs past the reference's 2048-entry table: same formula, longer
        eng.load_tensor(_lib.MODEL_TOKENIZER, "quantizer.pre_module.freqs_cis_stream", rope_table(8192 + 1024))
        eng.finalize(_lib.MODEL_TOKENIZER)
        unexpected = [k for k in unexpected if not k.startswith(("head.", "quantizer.post_module.", "quantizer.pre_module.",
                                                                 "quantizer.residual_bsq."))]
        return _IncompatibleKeys([], unexpected)

    @torch.no_grad()
    def encode(self, audios, audio_lengths):
        """FireflyArchitecture.encode, firefly_encoder.py:553-566: wav [B,L] f32, lens [B] ->
        (ids int64 [1,B,T], feature_lengths [B]).  Full-length rows go through the engine side by side in one call;
        ragged rows are encoded one by one: for a row shorter than L the causal-prefix property makes
        ids[:len//2048] identical to the reference, later ids are 0."""
        if torch.compiler.is_compiling():
            # under the caller's torch.compile(fullgraph=True) (infer_arvc.py:136-142) the call is one opaque custom op;
            # lengths cannot be read on the host there, rows are taken as full-length (the streaming loop's case)
            from . import ops
            return ops.enc_encode(audios), (audio_lengths // 512) // self.downsample_factor
        audios = audios.float()
        B, L = audios.shape
        dev = audios.device if audios.is_cuda else torch.device("cuda", self._engine.device)
        T = _lib.load().svanon_enc_num_ids(L)
        ids = torch.zeros(1, B, T, dtype=torch.int64, device=dev)
        lens = [int(x) for x in audio_lengths.reshape(-1).tolist()]
        if B > 1 and T > 0 and all(n >= L for n in lens):
            rows = audios.contiguous()
            _lib.check(self._engine.lib.svanon_enc_encode_batch(self._engine.handle, ptr(rows), B, L, ptr(ids),
                                                                C.c_void_p(_cuda_stream_ptr())))
            return ids, (audio_lengths // 512) // self.downsample_factor
        for b in range(B):
            n = min(lens[b], L)
            row = audios[b, :n].contiguous()
            tb = _lib.load().svanon_enc_num_ids(n)
            if tb == 0:
                continue
            out = ids[0, b, :tb] if tb == T else torch.empty(tb, dtype=torch.int64, device=dev)
            _lib.check(self._engine.lib.svanon_enc_encode(self._engine.handle, ptr(row), n, ptr(out),
                                                          C.c_void_p(_cuda_stream_ptr())))
            if tb != T:
                ids[0, b, :tb] = out
        feature_lengths = (audio_lengths // 512) // self.downsample_factor
        return ids, feature_lengths


class _Quantizer:
    def __init__(self, owner):
        self._o = owner
        self.downsample_factor = (2, 2)

    @torch.no_grad()
    def decode(self, indices):
        """DownsampleFiniteScalarQuantize.decode, fsq.py:112-116: [B,8,T] -> [B,512,4T] (a transposed view of the
        engine's channels-last buffer, which `head` consumes without a copy)."""
        if torch.compiler.is_compiling():
            from . import ops
            return ops.voc_quantizer_decode(indices).transpose(1, 2)
        eng = self._o._engine
        B, G, T = indices.shape
        assert G == 8
        dev = indices.device if indices.is_cuda else torch.device("cuda", eng.device)
        z = torch.empty(B, 4 * T, 512, dtype=torch.float32, device=dev)
        for b in range(B):
            codes = indices[b].to(torch.int64).contiguous()
            _lib.check(eng.lib.svanon_voc_quantizer_decode(eng.handle, ptr(codes), T, ptr(z[b]),
                                                           C.c_void_p(_cuda_stream_ptr())))
        return z.transpose(1, 2)


class _Head:
    def __init__(self, owner):
        self._o = owner

    @torch.no_grad()
    def __call__(self, z, template=None):
        """HiFiGANGenerator.forward, firefly.py:280-293: [B,512,L] -> [B,1,512 L]."""
        if torch.compiler.is_compiling():           # torch.compile(self.firefly.head, fullgraph=True), infer_arvc.py:128-134
            from . import ops
            return ops.voc_head(z)
        eng = self._o._engine
        B, Cc, L = z.shape
        assert Cc == 512
        zl = z.transpose(1, 2)
        if not zl.is_contiguous() or zl.dtype != torch.float32:
            zl = zl.float().contiguous()
        dev = zl.device if zl.is_cuda else torch.device("cuda", eng.device)
        wave = torch.empty(B, 1, 512 * L, dtype=torch.float32, device=dev)
        for b in range(B):
            _lib.check(eng.lib.svanon_voc_head(eng.handle, ptr(zl[b]), L, ptr(wave[b]), C.c_void_p(_cuda_stream_ptr())))
        return wave

    forward = __call__


class Vocoder(_Shim):
    _WANTED = ("head.", "quantizer.upsample.", "quantizer.residual_fsq.rvqs.", "backbone.", "quantizer.downsample.")

    def __init__(self, backbone=None, head=None, quantizer=None, spec_transform=None, device=None):
        self._engine = Engine.get(device)
        self.quantizer = _Quantizer(self)
        self.head = _Head(self)          # callers may rebind it (torch.compile, infer_arvc.py:128-134)
        self.downsample_factor = 4

    def load_state_dict(self, sd, strict: bool = False):
        eng = self._engine

        unexpected = eng.load_state_dict(_lib.MODEL_VOCODER, sd, lambda k: k.startswith(self._WANTED))
        # the encode path (reference wave -> codec ids of the prompt) is optional: decode-only checkpoints stay valid
        self.has_encoder = "backbone.norm.weight" in sd and "quantizer.residual_fsq.rvqs.0.project_in.weight" in sd
        if self.has_encoder:
            eng.load_tensor(_lib.MODEL_VOCODER, "spec_transform.fb", slaney_fbanks())
        eng.finalize(_lib.MODEL_VOCODER)         # folds weight norm (remove_parametrizations, infer_arvc.py:94)
        return _IncompatibleKeys([], unexpected)

    def remove_parametrizations(self):
        """firefly.py:597-602 -- weight norm is folded when the weights are finalized."""
        return None

    @torch.no_grad()
    def encode(self, audios, audio_lengths):
        """FireflyArchitecture.encode, firefly.py:561-574 (`wav2target_fn`, infer_arvc.py:168-171): wav [B,L] f32, lens [B]
        -> ((codes int32 [B,8,T], quantized), feature_lengths).  `quantized` (the FSQ latents, unused by the callers) is
        returned as None.  Rows shorter than L are encoded one by one on their valid prefix (causal-prefix property),
        later ids are 0."""
        if not getattr(self, "has_encoder", False):
            raise RuntimeError("the vocoder checkpoint was loaded without its encoder tensors (backbone.*, "
                               "quantizer.downsample.*, quantizer.residual_fsq.rvqs.*.project_in.*)")
        eng = self._engine
        audios = audios.float()
        B, L = audios.shape
        dev = audios.device if audios.is_cuda else torch.device("cuda", eng.device)
        T = L // 2048
        codes = torch.zeros(B, 8, T, dtype=torch.int32, device=dev)
        lens = [min(int(x), L) for x in audio_lengths.reshape(-1).tolist()]
        if T > 0 and all(n >= L for n in lens):
            rows = audios.contiguous()
            _lib.check(eng.lib.svanon_voc_encode(eng.handle, ptr(rows), B, L, ptr(codes), C.c_void_p(_cuda_stream_ptr())))
        else:
            for b in range(B):
                tb = lens[b] // 2048
                if tb == 0:
                    continue
                row = audios[b, : lens[b]].contiguous()
                out = torch.empty(8, tb, dtype=torch.int32, device=dev)
                _lib.check(eng.lib.svanon_voc_encode(eng.handle, ptr(row), 1, lens[b], ptr(out), C.c_void_p(_cuda_stream_ptr())))
                codes[b, :, :tb] = out
        return (codes, None), (audio_lengths // 512) // self.downsample_factor

    @torch.no_grad()
    def decode_codes(self, codes):
        """code2wav_fn (evaluations/infer_arvc.py:173-176) in one library call: [B,8,T] -> [B,1,2048 T]."""
        eng = self._engine
        B, G, T = codes.shape
        dev = codes.device if codes.is_cuda else torch.device("cuda", eng.device)
        wave = torch.empty(B, 1, 2048 * T, dtype=torch.float32, device=dev)
        for b in range(B):
            c = codes[b].to(torch.int64).contiguous()
            _lib.check(eng.lib.svanon_voc_decode(eng.handle, ptr(c), T, ptr(wave[b]), C.c_void_p(_cuda_stream_ptr())))
        return wave


class EncoderStream:
    """Stateful content encoder for `n_streams` streams side by side (SURVEY section 8b `enc_push_chunk`): feed each stream's
    NEW samples, get the ids of its new content frames.  Ids equal the reference's OFFLINE
    `FireflyArchitecture.encode()` (firefly_encoder.py:553-566) of the stream so far -- not the streaming loop's 128-frame
    window re-encode, which restarts from zero padding every chunk (SURVEY finding 4).

        es = EncoderStream(tokenizer, n_streams=4)
        ids = es.push(wave)            # wave [4, k * 2048] (k = 1..8 frames), host or device -> int64 [4, k]
        es.reset()                     # new utterances
    """

    def __init__(self, tokenizer: ContentTokenizer, n_streams: int = 1):
        self._engine = tokenizer._engine
        if not self._engine.loaded[_lib.MODEL_TOKENIZER]:
            raise RuntimeError("load the tokenizer weights first (ContentTokenizer.load_state_dict)")
        h = C.c_void_p()
        _lib.check(self._engine.lib.svanon_enc_stream_create(self._engine.handle, int(n_streams), C.byref(h)))
        self._h, self.n_streams = h, int(n_streams)

    @property
    def position(self) -> int:
        return int(self._engine.lib.svanon_enc_stream_position(self._h))

    @torch.no_grad()
    def push(self, wave: torch.Tensor) -> torch.Tensor:
        w = wave.reshape(self.n_streams, -1).float().contiguous()
        if w.shape[1] % 2048 or not 1 <= w.shape[1] // 2048 <= 8:
            raise ValueError("a push holds 1..8 whole frames of 2048 samples per stream")
        ids = torch.empty(self.n_streams, w.shape[1] // 2048, dtype=torch.int64, device=w.device)
        _lib.check(self._engine.lib.svanon_enc_push_chunk(self._h, ptr(w), w.shape[1], ptr(ids), C.c_void_p(_cuda_stream_ptr())))
        return ids

    def reset(self):
        _lib.check(self._engine.lib.svanon_enc_stream_reset(self._h, C.c_void_p(_cuda_stream_ptr())))

    def close(self):
        if self._h is not None:
            self._engine.lib.svanon_enc_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VocoderStream:
    """Stateful vocoder for `n_streams` streams side by side (SURVEY section 8b `voc_push_frames`): feed `frames_per_push` new
    code frames per stream, get their samples.  Pushing an utterance frame by frame reproduces `Vocoder.head(quantizer.decode(
    codes))` of the whole utterance (every conv of firefly.py:280-293 is causal; per-layer history is kept on the device).

        vs = VocoderStream(vocoder, n_streams=2, frames_per_push=1)
        wave = vs.push(codes)          # codes [2, 8, 1] int -> float32 [2, 2048]
    """

    def __init__(self, vocoder: Vocoder, n_streams: int = 1, frames_per_push: int = 1):
        self._engine = vocoder._engine
        if not self._engine.loaded[_lib.MODEL_VOCODER]:
            raise RuntimeError("load the vocoder weights first (Vocoder.load_state_dict)")
        h = C.c_void_p()
        _lib.check(self._engine.lib.svanon_voc_stream_create(self._engine.handle, int(n_streams), int(frames_per_push), C.byref(h)))
        self._h, self.n_streams, self.frames = h, int(n_streams), int(frames_per_push)

    @torch.no_grad()
    def push(self, codes: torch.Tensor) -> torch.Tensor:
        c = codes.reshape(self.n_streams, 8, self.frames).to(torch.int64).contiguous()
        dev = c.device
        out = torch.empty(self.n_streams, self.frames * 2048, dtype=torch.float32, device=dev)
        _lib.check(self._engine.lib.svanon_voc_push_frames(self._h, ptr(c), ptr(out), C.c_void_p(_cuda_stream_ptr())))
        return out

    def reset(self):
        _lib.check(self._engine.lib.svanon_voc_stream_reset(self._h, C.c_void_p(_cuda_stream_ptr())))

    def close(self):
        if self._h is not None:
            self._engine.lib.svanon_voc_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

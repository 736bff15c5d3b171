"""Per-chunk streaming loop oracle (E -> A -> V).  TEST INFRASTRUCTURE ONLY.

Restates `InferenceWrapper.setup_stream_caches`, `prefill_prompt` (the part after the
speaker encoders) and `process_one_chunk` (evaluations/infer_arvc.py:443-596) on top of
the three stage oracles.  The two speaker-encoder outputs (style_vectors,
timbre_latents) are inputs: those encoders are on the setup path.
"""
from __future__ import annotations

import torch

from . import content_encoder as E
from . import vocoder as V
from .dual_ar import DualAR

SAMPLES_PER_FRAME = 2048
NUM_CODEBOOKS = 8


class StreamOracle:
    def __init__(self, ar_sd, tok_sd, voc_sd_folded, noise_fn):
        self.ar = DualAR(ar_sd, noise_fn)
        self.tok_sd = tok_sd
        self.voc_sd = voc_sd_folded
        self.timings = {"E": [], "A": [], "V": []}

    # infer_arvc.py:463-489 with calculate_prompt's encoder outputs supplied by the caller
    def prefill_prompt(self, ref_audio_codes, ref_content_codes, style_vectors, timbre_latents,
                       max_prompt_frames=256, delay=2):
        # NOTE (reference behaviour): the truncated copies are kept for window padding and
        # re-prompting, but the UNtruncated codes are what is prefilled (:484-489).
        self.ref_audio_codes = ref_audio_codes[:, :, :max_prompt_frames]
        self.ref_content_codes = ref_content_codes[:, :max_prompt_frames]
        self.style_vectors = style_vectors
        self.timbre_latents = timbre_latents
        self.delay = int(delay)
        self.ar.set_delay(delay)
        self.ar.prefill_prompt(ref_content_codes, ref_audio_codes, style_vectors, timbre_latents)

    # infer_arvc.py:443-460
    def setup_stream_caches(self, encode_window_frames=128, decode_window_frames=64, max_seq_frames=768,
                            buffer_frames=32, decode_chunk_frames=1):
        self.src_wav = torch.zeros(1, encode_window_frames * SAMPLES_PER_FRAME)
        self.encode_window_frames = encode_window_frames
        self.decode_window_frames = decode_window_frames
        self.max_seq_frames = max_seq_frames
        self.buffer_frames = buffer_frames
        self.chunk = decode_chunk_frames
        self.src_content_codes = torch.zeros(1, 0, dtype=torch.long)
        self.pred_codes = torch.zeros(1, NUM_CODEBOOKS, 0, dtype=torch.long)
        self.prefilled = False

    # infer_arvc.py:492-596
    def process_one_chunk(self, chunk):
        import time
        n = chunk.size(-1)
        self.src_wav[:, :-n] = self.src_wav[:, n:].clone()
        self.src_wav[:, -n:] = chunk
        t0 = time.perf_counter()
        ids = E.encode(self.src_wav, self.tok_sd)[0].squeeze(0)
        self.timings["E"].append(time.perf_counter() - t0)
        self.src_content_codes = torch.cat([self.src_content_codes, ids[..., -self.chunk:]], dim=-1)
        if self.src_content_codes.size(-1) < self.delay:
            return torch.zeros_like(chunk)
        if not self.prefilled and self.delay != 0:
            self.ar.prefill_src_condition4delay(self.src_content_codes[:, -self.delay:])
            self.prefilled = True
            return torch.zeros_like(chunk)
        t0 = time.perf_counter()
        for i in range(self.chunk):
            code, pos = self.ar.decode_one(ids[..., -(self.chunk - i)].unsqueeze(0))
            self.pred_codes = torch.cat([self.pred_codes, code.clone()[None].long()], dim=-1)
        self.timings["A"].append(time.perf_counter() - t0)
        if int(pos) // 2 >= self.max_seq_frames:
            ext_audio = torch.cat([self.ref_audio_codes.long(), self.pred_codes[..., -self.buffer_frames:]], dim=-1)
            ext_content = torch.cat([self.ref_content_codes,
                                     self.src_content_codes[..., -self.buffer_frames - self.delay:-self.delay]], dim=-1)
            self.ar.prefill_prompt(ext_content, ext_audio, self.style_vectors, self.timbre_latents)
            self.ar.prefill_src_condition4delay(self.src_content_codes[..., -self.delay:])
        win = self.pred_codes[..., -self.decode_window_frames:]
        pad = self.decode_window_frames - win.size(-1)
        if pad > 0:
            win = torch.cat([self.ref_audio_codes[..., -pad:].long(), win], dim=-1)
        t0 = time.perf_counter()
        wave = V.code2wav(win.reshape(1, NUM_CODEBOOKS, self.decode_window_frames), self.voc_sd)
        self.timings["V"].append(time.perf_counter() - t0)
        self.pred_codes = self.pred_codes[..., -SAMPLES_PER_FRAME:]
        self.src_content_codes = self.src_content_codes[..., -SAMPLES_PER_FRAME:]
        return wave[..., -SAMPLES_PER_FRAME * self.chunk:].squeeze(1)

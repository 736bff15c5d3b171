"""`evaluations.infer_arvc.InferenceWrapper` over the engine -- the caller of the hot path, method by method:

    calculate_prompt(ref_wav_tensors, alpha, spk_emb_collate_type)          infer_arvc.py:382-441
    prefill_prompt(ref_wav_tensors, max_prompt_frames, delay, alpha, ...)   :462-489
    setup_stream_caches(encode_window_frames, decode_window_frames, ...)    :443-460
    process_one_chunk(src_wav_chunk [1, chunk * 2048]) -> [1, chunk * 2048]  :492-596
    stream_infer(src, refs, ...) -> np.ndarray                              :598-676
    infer(src, refs, delay=, alpha=) -> np.ndarray                          :261-380

Same method names, argument names, defaults and return shapes as the reference; what differs is on the outside only:
sources and references are passed as tensors (or .wav paths read with scipy; `librosa` / `torchaudio.save` are absent
offline), and results are never written to disk.  Every numeric step is a library call of libsvanon_b200.so: the
streaming loop is `svanon_stream_process_chunk` (one call per chunk), the prompt is `PromptBuilder.calculate_prompt`.
The two `torch.randn_like` draws of the anonymisation mix come from torch's global generator exactly where the
reference takes them (style first), so a shared `torch.manual_seed` reproduces the reference's mix."""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Optional, Sequence, Union

import numpy as np
import torch

from .prompt import PromptBuilder
from .streaming import StreamSession, new_sampler_seed

Wave = Union[str, Path, torch.Tensor, np.ndarray]


class InferenceWrapper:
    SAMPLES_PER_FRAME = 2048
    NUM_CODEBOOKS = 8
    RESAMPLE_FREQ = 16000
    MEL_BINS = 80

    def __init__(self, model, speech_tokenizer, firefly, style_encoder, timbre_encoder, sr: int = 44100, device=None):
        """The five model objects are the engine's shims (ARVCWrapper, ContentTokenizer, Vocoder, speaker.CAMPPlus,
        speaker.SpeakerEncoder) with their weights loaded -- what `InferenceWrapper.__init__` builds from the YAMLs and
        checkpoints (infer_arvc.py:33-126)."""
        self.model, self.speech_tokenizer, self.firefly = model, speech_tokenizer, firefly
        self.style_encoder, self.timbre_encoder = style_encoder, timbre_encoder
        self.sr = sr
        self.device = torch.device(device) if device is not None else torch.device("cuda", model._engine.device)
        self._prompt = PromptBuilder(speech_tokenizer, firefly, style_encoder, timbre_encoder, sr, self.RESAMPLE_FREQ)
        self._session: Optional[StreamSession] = None
        self._noise_fn: Optional[Callable] = None

    @classmethod
    def from_state_dicts(cls, ar_sd, tokenizer_sd, vocoder_sd, style_sd, timbre_sd, max_seq_len: int = 2048, device=None):
        """Builds and loads the five shims from reference-keyed state dicts (the five checkpoint paths of
        configs/config_firefly_arvcasr_8192_delay0_8.yaml:43-57)."""
        from . import ARVCWrapper, ContentTokenizer, Vocoder
        from .speaker import CAMPPlus, SpeakerEncoder
        model = ARVCWrapper()
        model.setup_caches(max_batch_size=1, max_seq_len=max_seq_len, dtype=torch.float16)
        model.load_state_dict(ar_sd, strict=False)
        tok = ContentTokenizer()
        tok.load_state_dict(tokenizer_sd, strict=False)
        voc = Vocoder()
        voc.load_state_dict(vocoder_sd, strict=False)
        voc.remove_parametrizations()
        style = CAMPPlus()
        style.load_state_dict(style_sd, strict=False)
        timbre = SpeakerEncoder()
        timbre.load_state_dict(timbre_sd, strict=False)
        return cls(model, tok, voc, style, timbre, device=device)

    @classmethod
    def from_config(cls, config_path, checkpoint_path, compile_encoder=False, compile_decoder=False, compile_ar=False,
                    fp16=False, root=None):
        """The reference's constructor (`InferenceWrapper(config_path, checkpoint_path, compile_*=...)`,
        evaluations/infer_arvc.py:33-126; `load_models` of real-time-gui.py:52-57): reads the top-level YAML, loads the AR
        checkpoint and the four helper checkpoints it names (`speech_tokenizer` / `firefly` / `style_encoder` /
        `timbre_encoder` -> `checkpoint_path`; the tokenizer may be wrapped in {'net': ...} with a 'module.' prefix,
        :73-78) and builds the engine's shims from them.  The `compile_*` / `fp16` switches are accepted and have nothing
        to do: the decode step already is one kernel, encode / head are library calls, the parity build is fp32.
        Relative paths resolve against `root` (default: the working directory, like the reference CLI)."""
        import yaml
        base = Path(root) if root is not None else Path.cwd()

        def resolve(p):
            p = Path(p)
            return p if p.is_absolute() else base / p

        config = yaml.safe_load(open(resolve(config_path)))

        def ckpt(section):
            return torch.load(resolve(config[section]["checkpoint_path"]), map_location="cpu")

        tok_sd = ckpt("speech_tokenizer")
        if "net" in tok_sd:
            tok_sd = tok_sd["net"]
        tok_sd = {k[7:] if k.startswith("module.") else k: v for k, v in tok_sd.items()}
        iw = cls.from_state_dicts(torch.load(resolve(checkpoint_path), map_location="cpu"), tok_sd, ckpt("firefly"),
                                  ckpt("style_encoder"), ckpt("timbre_encoder"))
        iw.config = config
        iw.sr = config["preprocess_params"]["sr"]
        iw._prompt = PromptBuilder(iw.speech_tokenizer, iw.firefly, iw.style_encoder, iw.timbre_encoder, iw.sr, cls.RESAMPLE_FREQ)
        return iw

    # ------------------------------------------------------------------------------------------ helpers
    def set_noise_fn(self, fn: Optional[Callable]):
        """Sampling-noise tape `fn(step, slot, V)` shared with the oracle (tests); None: the library's own generator."""
        self._noise_fn = fn

    def _load(self, wave: Wave, crop_seconds: Optional[float] = None) -> torch.Tensor:
        """`librosa.load(path, sr=self.sr)` (infer_arvc.py:274,615,623) for 16-bit / float .wav files, or a tensor /
        array that already is at self.sr  ->  [1, n] float32 on the device.  A file at another rate is converted with the
        engine's windowed-sinc resampler (torchaudio semantics); librosa's default is soxr_hq, so such files give close
        but not bit-identical samples -- parity claims are made on audio that already is at the model rate."""
        if isinstance(wave, (str, Path)):
            from scipy.io import wavfile
            rate, data = wavfile.read(str(wave))
            data = np.asarray(data)
            if data.dtype == np.uint8:                               # 8-bit PCM is unsigned, offset 128
                x = (torch.from_numpy(data.astype(np.float32)) - 128.0) / 128.0
            elif np.issubdtype(data.dtype, np.integer):              # int16, or int32 for 24/32-bit PCM: full scale of the dtype
                x = torch.from_numpy(data.astype(np.float64) / (float(np.iinfo(data.dtype).max) + 1.0)).float()
            elif np.issubdtype(data.dtype, np.floating):
                x = torch.from_numpy(data.astype(np.float32))
            else:
                raise ValueError(f"unsupported .wav sample type {data.dtype}")
            if x.dim() == 2:
                x = x.mean(dim=1)                                    # librosa mono=True: channel mean
            x = x[None].to(self.device)
            if rate != self.sr:
                from .audio import Resampler
                x = Resampler(rate, self.sr)(x)
        else:
            x = torch.as_tensor(wave, dtype=torch.float32).reshape(1, -1).to(self.device)
        if crop_seconds is not None:
            x = x[:, : int(crop_seconds * self.sr)]
        return x.contiguous()

    def process_ref_paths(self, ref_path, ref_crop_lengths):
        """infer_arvc.py:234-247: one reference or a list; one optional crop length (seconds) per reference."""
        refs = list(ref_path) if isinstance(ref_path, (list, tuple)) else [ref_path]
        if ref_crop_lengths is None:
            crops = [None] * len(refs)
        elif isinstance(ref_crop_lengths, (int, float)):
            crops = [ref_crop_lengths] * len(refs)
        else:
            crops = list(ref_crop_lengths)
            if len(crops) != len(refs):
                raise ValueError("ref_crop_lengths must have one entry per reference")
        return refs, crops

    def create_wave_lens_tensor(self, tensor):
        return torch.LongTensor([tensor.size(-1)])

    def apply_noise_mixing(self, tensor, alpha):
        from .prompt import apply_noise_mixing
        return apply_noise_mixing(tensor, alpha)

    def code2wav_fn(self, code):
        """infer_arvc.py:173-176."""
        return self.firefly.head(self.firefly.quantizer.decode(code))

    # ------------------------------------------------------------------------------------------ prompt
    @torch.no_grad()
    def calculate_prompt(self, ref_wav_tensors, alpha=1.0, spk_emb_collate_type="concat_mel", *, noise_style=None,
                         noise_timbre=None):
        """`noise_style` / `noise_timbre` (extension, keyword-only): the standard-normal draws of the anonymisation mix,
        to replay a recorded run; default: `torch.randn_like` from the global generator, as the reference."""
        return self._prompt.calculate_prompt(ref_wav_tensors, alpha, spk_emb_collate_type, noise_style, noise_timbre)

    @torch.no_grad()
    def prefill_prompt(self, ref_wav_tensors, max_prompt_frames=256, delay=4, alpha=1.0, spk_emb_collate_type="concat_mel",
                       *, noise_style=None, noise_timbre=None):
        codes, content, style, timbre, ref = self.calculate_prompt(ref_wav_tensors, alpha, spk_emb_collate_type,
                                                                   noise_style=noise_style, noise_timbre=noise_timbre)
        # the reference keeps truncated copies for padding / re-prompting and prefills the UNtruncated prompt
        # (infer_arvc.py:469-489); the stream session does both from the full tensors
        self.ref_audio_codes = codes[:, :, :max_prompt_frames]
        self.ref_content_codes = content[:, :max_prompt_frames]
        self.style_vectors, self.timbre_latents = style, timbre
        self.ref_wav_tensor = ref[:, : max_prompt_frames * self.SAMPLES_PER_FRAME]
        self.ref_wav_tensor_len = self.ref_wav_tensor.size(-1)
        original = self.model.decoder.original_delay
        self.delay = original if isinstance(original, int) else int(delay)
        self.model.set_delay(delay=delay)
        if self._session is not None:
            self._session.close()
        self._session = StreamSession(self.device, getattr(self.model, "_max_seq_len", 2048))
        if self._noise_fn is not None:           # else: the new session seeded the library's generator from torch's
            self._session.set_noise_fn(self._noise_fn, 0)
        self._session.set_prompt(content[0], codes, style, timbre, max_prompt_frames=max_prompt_frames, delay=self.delay)

    def setup_stream_caches(self, encode_window_frames=96, decode_window_frames=64, max_seq_frames=768, buffer_frames=32,
                            decode_chunk_frames=1, delay=None):
        if self._session is None:
            raise RuntimeError("prefill_prompt first (the reference's stream_infer order, infer_arvc.py:632-646)")
        self.encode_window_frames, self.decode_window_frames = encode_window_frames, decode_window_frames
        self.max_seq_frames, self.buffer_frames, self.decode_chunk_frames = max_seq_frames, buffer_frames, decode_chunk_frames
        self._session.setup(encode_window_frames, decode_window_frames, max_seq_frames, buffer_frames, decode_chunk_frames)

    # ------------------------------------------------------------------------------------------ the loop
    @torch.no_grad()
    def process_one_chunk(self, src_wav_chunk: torch.Tensor, pitch_shift: float = 0.0) -> torch.Tensor:
        """infer_arvc.py:492-596: [1, chunk * 2048] -> [1, chunk * 2048] (zeros while the delay warm-up fills).
        `pitch_shift` is accepted and unused, as in the reference (:494 -- nothing in its body reads it)."""
        n = self.decode_chunk_frames * self.SAMPLES_PER_FRAME
        if src_wav_chunk.shape[-1] != n:
            raise ValueError(f"chunk of {src_wav_chunk.shape[-1]} samples, expected decode_chunk_frames * 2048 = {n}")
        return self._session.process_chunk(src_wav_chunk.reshape(-1))[None]

    @property
    def src_content_codes(self) -> torch.Tensor:
        return self._session.history()[0][None]

    @property
    def pred_codes(self) -> torch.Tensor:
        return self._session.history()[1][None]

    @torch.no_grad()
    def stream_infer(self, src_path: Wave, ref_path: Union[Wave, Sequence[Wave]], out_dir=None, encode_window_frames=128,
                     decode_window_frames=64, max_prompt_frames=256, max_seq_frames=768, buffer_frames=32,
                     decode_chunk_frames=1, delay=None, ref_crop_lengths=None, alpha=1.0,
                     spk_emb_collate_type="concat_mel", save_result=False) -> np.ndarray:
        if save_result:
            raise NotImplementedError("writing .wav files is left to the caller (torchaudio.save needs torchcodec)")
        src = self._load(src_path)
        refs, crops = self.process_ref_paths(ref_path, ref_crop_lengths)
        ref_tensors = [self._load(r, c) for r, c in zip(refs, crops)]
        self.prefill_prompt(ref_tensors, max_prompt_frames=max_prompt_frames, delay=delay, alpha=alpha,
                            spk_emb_collate_type=spk_emb_collate_type)
        self.setup_stream_caches(encode_window_frames, decode_window_frames, max_seq_frames, buffer_frames, decode_chunk_frames)
        step = self.SAMPLES_PER_FRAME * decode_chunk_frames
        pad = step - src.size(1) % step            # a whole extra chunk when already aligned (infer_arvc.py:648-649)
        src = torch.nn.functional.pad(src, (pad, 0), value=0)
        chunks = src.view(-1, step)
        out = torch.empty_like(chunks)
        for i in range(chunks.shape[0]):
            self._session.process_chunk(chunks[i], out[i])
        return out.reshape(-1).cpu().numpy()

    # ------------------------------------------------------------------------------------------ offline
    @torch.no_grad()
    def infer(self, src_path: Wave, ref_path: Union[Wave, Sequence[Wave]], out_dir=None, output_path=None, delay=None,
              ref_crop_lengths=None, alpha=1.0, spk_emb_collate_type="concat_mel", save_result=False, *,
              noise_style=None, noise_timbre=None, **sampling_kwargs) -> np.ndarray:
        if save_result:
            raise NotImplementedError("writing .wav files is left to the caller (torchaudio.save needs torchcodec)")
        src = self._load(src_path)
        refs, crops = self.process_ref_paths(ref_path, ref_crop_lengths)
        ref_tensors = [self._load(r, c) for r, c in zip(refs, crops)]
        # same arithmetic and the same two noise draws as the reference's inlined copy of calculate_prompt (:280-346)
        codes, content, style, timbre, _ = self._prompt.calculate_prompt(ref_tensors, alpha, spk_emb_collate_type, noise_style,
                                                                         noise_timbre, allow_avg=True)
        src_content, _ = self.speech_tokenizer.encode(src, self.create_wave_lens_tensor(src))
        if delay is not None:
            self.model.set_delay(delay=delay)
        self.model.set_noise_fn(self._noise_fn, 0)          # None too: a stale tape must not be replayed
        if self._noise_fn is None:
            self.model.set_sampling(seed=new_sampler_seed())
        vc_codes = self.model.generate(ref_content_codes=content, ref_audio_codes=codes, src_content_codes=src_content.squeeze(0),
                                       style_vectors=style, timbre_latents=timbre, **sampling_kwargs)
        return self.code2wav_fn(vc_codes.long()).squeeze().cpu().numpy()

    # ------------------------------------------------------------------------------------------ offline, many utterances
    @torch.no_grad()
    def infer_batch(self, src_paths: Sequence[Wave], ref_paths: Sequence[Union[Wave, Sequence[Wave]]], delay: Optional[int] = None,
                    alpha: float = 1.0, spk_emb_collate_type: str = "concat_mel", noise_fns: Optional[Sequence[Callable]] = None,
                    noises_style=None, noises_timbre=None):
        """BASELINE config 3: `infer` for many (source, reference) pairs at once.  The reference is batch-1 and would call
        `infer` once per pair (evaluations/infer_arvc.py:56,261-380); here the prompts are built per pair, and the
        autoregressive decode of ALL pairs advances in lock-step with one pass over the weights per frame
        (svanon_ar_generate_many; pairs of different length leave the batch as they finish).  Pair k yields exactly what
        `infer(src_paths[k], ref_paths[k], delay=delay, alpha=alpha)` yields.  Returns a list of waveforms (np.ndarray).

        `noise_fns[k]` (tests): sampling-noise tape of pair k; `noises_style[k]` / `noises_timbre[k]`: recorded draws of
        the anonymisation mix (default: torch's global generator, style then timbre, pair after pair)."""
        import ctypes as C
        from . import _lib
        from .engine import ptr, _cuda_stream_ptr
        n = len(src_paths)
        if n == 0 or len(ref_paths) != n:
            raise ValueError("one reference (or list of references) per source")
        if delay is None:                               # `infer`'s default: whatever the model is set to
            delay = self.model.decoder.delay if isinstance(self.model.decoder.delay, int) else 0
        eng = self.model._engine
        lib = eng.lib
        keep, streams = [], []
        rc_p, ra_p, sc_p, sv_p, tl_p, nz_p, out_p = ([] for _ in range(7))
        Tr, Ts, outs = [], [], []
        try:
            for k in range(n):
                src = self._load(src_paths[k])
                refs, crops = self.process_ref_paths(ref_paths[k], None)
                ref_tensors = [self._load(r, c) for r, c in zip(refs, crops)]
                codes, content, style, timbre, _ = self._prompt.calculate_prompt(
                    ref_tensors, alpha, spk_emb_collate_type, None if noises_style is None else noises_style[k],
                    None if noises_timbre is None else noises_timbre[k], allow_avg=True)
                src_content, _ = self.speech_tokenizer.encode(src, self.create_wave_lens_tensor(src))
                rc = content.reshape(-1).to(self.device, torch.int64).contiguous()
                ra = codes.reshape(8, -1).to(self.device, torch.int32).contiguous()
                sc = src_content.reshape(-1).to(self.device, torch.int64).contiguous()
                sv = style.reshape(-1).to(self.device, torch.float32).contiguous()
                tl = timbre.reshape(32, 128).to(self.device, torch.float32).contiguous()
                out = torch.empty(8, sc.numel(), dtype=torch.int32, device=self.device)
                noise = None
                if noise_fns is not None and noise_fns[k] is not None:
                    noise = torch.stack([torch.stack([noise_fns[k](i, s, 1000)[:1000] for s in range(1, 9)])
                                         for i in range(sc.numel())]).float().contiguous().to(self.device)
                h = C.c_void_p()
                _lib.check(lib.svanon_stream_create(eng.handle, getattr(self.model, "_max_seq_len", 2048), C.byref(h)))
                streams.append(h)
                _lib.check(lib.svanon_ar_set_delay(h, int(delay)))
                if noise is None:                      # the library's counter-based generator: a fresh seed per utterance,
                    _lib.check(lib.svanon_ar_set_sampling(h, 0.7, 0.7, new_sampler_seed()))   # drawn from torch's generator
                keep += [rc, ra, sc, sv, tl, out, noise]
                for lst, t in ((rc_p, rc), (ra_p, ra), (sc_p, sc), (sv_p, sv), (tl_p, tl), (out_p, out)):
                    lst.append(t.data_ptr())
                nz_p.append(noise.data_ptr() if noise is not None else None)
                Tr.append(rc.numel())
                Ts.append(sc.numel())
                outs.append(out)
            arr = lambda ps: (C.c_void_p * n)(*ps)                                       # noqa: E731
            ints = lambda v: (C.c_int * n)(*v)                                           # noqa: E731
            _lib.check(lib.svanon_ar_generate_many(arr([s.value for s in streams]), n, arr(rc_p), arr(ra_p), ints(Tr), arr(sc_p),
                                                   ints(Ts), arr(sv_p), arr(tl_p), arr(nz_p), arr(out_p),
                                                   C.c_void_p(_cuda_stream_ptr())))
            waves = [self.code2wav_fn(o[None].long()).squeeze().cpu().numpy() for o in outs]
        finally:
            for h in streams:
                lib.svanon_stream_destroy(h)
        del keep
        return waves

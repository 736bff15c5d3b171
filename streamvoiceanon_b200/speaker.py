"""The two speaker encoders of the prompt path over the C ABI -- reference-facing mirrors of

    modules.campplus.DTDNN.CAMPPlus                      `style_encoder(feat [B,T,80], feat_lens [B]) -> [B,192]`
    modules.bicodec_speaker_encoder.speaker_encoder.SpeakerEncoder
                                                         `tokenize_wav(wav [B,n], wav_lens [B]) -> (zq [B,128,32], indices)`
    InferenceWrapper.calculate_style_vec / calculate_timbre_latent   (evaluations/infer_arvc.py:179-223)
    torchaudio.compliance.kaldi.fbank (num_mel_bins=80, dither=0, sample_frequency=16000; :186-191)

The arithmetic runs in libsvanon_b200.so (csrc/speaker.hpp, speaker.cu); there is no CPU path.  Four derived buffers
are built here with the reference's own torch formulas and uploaded with the checkpoint: the povey window and kaldi
mel banks of the fbank, the centred hann window and the slaney filterbank of the timbre encoder's MelSpectrogram."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import torch

from . import _lib
from .engine import Engine, ptr, _cuda_stream_ptr, slaney_fbanks

MODEL_STYLE, MODEL_TIMBRE = 3, 4


# ---------------------------------------------------------------------------------------------- derived buffers
def povey_window(win: int = 400) -> torch.Tensor:
    """torchaudio.compliance.kaldi `_feature_window_function("povey")`: hann(win, symmetric) ** 0.85."""
    return torch.hann_window(win, periodic=False).pow(0.85)


def kaldi_mel_banks(num_bins: int = 80, padded: int = 512, sr: float = 16000.0, low: float = 20.0) -> torch.Tensor:
    """torchaudio.compliance.kaldi `get_mel_banks` without VTLN warping: triangles linear in the mel domain between
    mel(low) and mel(nyquist), plus the zero Nyquist column kaldi.fbank pads  ->  [num_bins, padded/2 + 1]."""
    def mel(f):
        return 1127.0 * math.log(1.0 + f / 700.0)
    n_fft_bins = padded // 2
    mel_low, mel_high = mel(low), mel(0.5 * sr)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    m = (1127.0 * (1.0 + (sr / padded) * torch.arange(n_fft_bins) / 700.0).log()).unsqueeze(0)
    bins = torch.max(torch.zeros(1), torch.min((m - left) / (center - left), (right - m) / (right - center)))
    return torch.nn.functional.pad(bins, (0, 1), value=0.0).contiguous()


def centred_hann(n_fft: int = 1024, win_length: int = 640) -> torch.Tensor:
    """The window `torch.stft` applies for win_length < n_fft: hann(win_length, periodic) zero-padded on both sides."""
    w = torch.zeros(n_fft)
    lo = (n_fft - win_length) // 2
    w[lo: lo + win_length] = torch.hann_window(win_length)
    return w


def style_derived_buffers() -> Dict[str, torch.Tensor]:
    return {"fbank.window": povey_window(), "fbank.mel_banks": kaldi_mel_banks()}


def timbre_derived_buffers() -> Dict[str, torch.Tensor]:
    """`mel_fn` of configs/hydra_arcs/sv/sparktts_speaker_encoder.yaml: n_fft 1024, win 640, hop 320, 128 slaney mel bins
    from 10 Hz to 8 kHz at 16 kHz."""
    return {"mel.window": centred_hann(),
            "mel.fb": slaney_fbanks(n_freqs=513, f_min=10.0, f_max=8000.0, n_mels=128, sample_rate=16000).contiguous()}


def _style_key(k: str) -> str:
    """CAMPPlus.load_state_dict renames the checkpoint's `xvector.dense.*` to `dense.*` (DTDNN.py:114-130)."""
    return k.replace("xvector.dense", "dense") if k.startswith("xvector.dense") else k


# ---------------------------------------------------------------------------------------------- style
class CAMPPlus:
    """`modules.campplus.DTDNN.CAMPPlus(feat_dim=80, embedding_size=192)` (configs/hydra_arcs/sv/campplus.yaml)."""

    def __init__(self, feat_dim: int = 80, embedding_size: int = 192, device=None, **_unused):
        if feat_dim != 80 or embedding_size != 192:
            raise ValueError("the engine implements the shipped CAMPPlus configuration (feat_dim=80, embedding_size=192)")
        self._engine = Engine.get(device)

    def to(self, *_a, **_k):
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(())

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False) -> Tuple[list, list]:
        sd = {_style_key(k): v for k, v in sd.items()}
        unexpected = self._engine.load_state_dict(MODEL_STYLE, sd, lambda k: True)
        for k, v in style_derived_buffers().items():
            self._engine.load_tensor(MODEL_STYLE, k, v)
        self._engine.finalize(MODEL_STYLE)
        return [], unexpected

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, x_lens: torch.Tensor = None) -> torch.Tensor:
        """DTDNN.py:132-138: x [B, T, 80] features, x_lens [B] valid rows after the stride-2 TDNN -> [B, 192]."""
        if x.dim() != 3 or x.shape[2] != 80:
            raise ValueError(f"expected features [B, T, 80], got {tuple(x.shape)}")
        B, T = x.shape[0], x.shape[1]
        rows_after = (T - 1) // 2 + 1
        lens = [rows_after] * B if x_lens is None else [int(v) for v in x_lens.reshape(-1).tolist()]
        feat = x.detach().to(torch.float32).contiguous()
        out = torch.empty(B, 192, dtype=torch.float32, device=feat.device)
        for b in range(B):
            _lib.check(self._engine.lib.svanon_campplus_forward(self._engine.handle, ptr(feat[b]), T, lens[b], ptr(out[b]),
                                                                C.c_void_p(_cuda_stream_ptr())))
        return out

    forward = __call__


@torch.no_grad()
def kaldi_fbank(waveform: torch.Tensor, num_mel_bins: int = 80, dither: float = 0.0, sample_frequency: float = 16000.0,
                device=None) -> torch.Tensor:
    """`torchaudio.compliance.kaldi.fbank` for the one argument set the reference uses: waveform [1, n] -> [m, 80]."""
    if num_mel_bins != 80 or dither != 0.0 or sample_frequency != 16000.0:
        raise ValueError("only num_mel_bins=80, dither=0, sample_frequency=16000 (evaluations/infer_arvc.py:186-191)")
    eng = Engine.get(device if device is not None else (waveform.device if waveform.is_cuda else None))
    w = waveform.reshape(-1).detach().to(torch.float32).contiguous()
    n = w.numel()
    m = 0 if n < 400 else 1 + (n - 400) // 160
    out = torch.empty(m, 80, dtype=torch.float32, device=w.device)
    if m:
        _lib.check(eng.lib.svanon_kaldi_fbank(eng.handle, ptr(w), n, ptr(out), C.c_void_p(_cuda_stream_ptr())))
    return out


@torch.no_grad()
def calculate_style_vec(style_encoder: CAMPPlus, audio_16k_tensor: torch.Tensor, wave_lens: torch.Tensor) -> torch.Tensor:
    """`InferenceWrapper.calculate_style_vec` (evaluations/infer_arvc.py:179-211).  One row (what `calculate_prompt`
    passes): a single library call, wave -> [1, 192].  Several rows: per-row fbank, rows padded to the longest with the
    row's minimum and lens = frames // 2, as the reference does."""
    eng = style_encoder._engine
    B = audio_16k_tensor.shape[0]
    lens = [int(v) for v in wave_lens.reshape(-1).tolist()]
    if B == 1:
        w = audio_16k_tensor[0, : lens[0]].detach().to(torch.float32).contiguous()
        out = torch.empty(1, 192, dtype=torch.float32, device=w.device)
        _lib.check(eng.lib.svanon_style_vector(eng.handle, ptr(w), w.numel(), ptr(out), C.c_void_p(_cuda_stream_ptr())))
        return out
    feats = []
    for b in range(B):
        f = kaldi_fbank(audio_16k_tensor[b: b + 1, : lens[b]])
        feats.append(f - f.mean(dim=0, keepdim=True))
    longest = max(f.shape[0] for f in feats)
    feat_lens = torch.tensor([f.shape[0] for f in feats], dtype=torch.int32) // 2
    feats = [torch.nn.functional.pad(f, (0, 0, 0, longest - f.shape[0]), value=float(f.min().item())) for f in feats]
    return style_encoder(torch.stack(feats, dim=0), feat_lens)


# ---------------------------------------------------------------------------------------------- timbre
class SpeakerEncoder:
    """`modules.bicodec_speaker_encoder.speaker_encoder.SpeakerEncoder` as configured by
    configs/hydra_arcs/sv/sparktts_speaker_encoder.yaml -- the `tokenize_wav` path only (mel -> ECAPA-TDNN trunk ->
    PerceiverResampler -> FSQ 4^6)."""

    def __init__(self, device=None, **_unused):
        self._engine = Engine.get(device)

    def to(self, *_a, **_k):
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(())

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False) -> Tuple[list, list]:
        wanted = ("speaker_encoder.layer", "speaker_encoder.conv.", "perceiver_sampler.", "quantizer.project_")
        unexpected = self._engine.load_state_dict(MODEL_TIMBRE, sd, lambda k: k.startswith(wanted))
        for k, v in timbre_derived_buffers().items():
            self._engine.load_tensor(MODEL_TIMBRE, k, v)
        self._engine.finalize(MODEL_TIMBRE)
        return [], unexpected

    @torch.no_grad()
    def tokenize_wav(self, wav: torch.Tensor, wav_lens: torch.Tensor):
        """speaker_encoder.py:136-144: wav [B, n] at 16 kHz (rows zero-padded to n), wav_lens [B] -> (zq [B, 128, 32],
        indices [B, 1, 32] int32)."""
        B, n = wav.shape
        w = wav.detach().to(torch.float32).contiguous()
        lens = [int(v) for v in wav_lens.reshape(-1).tolist()]
        lat = torch.empty(B, 32, 128, dtype=torch.float32, device=w.device)
        idx = torch.empty(B, 1, 32, dtype=torch.int32, device=w.device)
        for b in range(B):
            _lib.check(self._engine.lib.svanon_timbre_latent(self._engine.handle, ptr(w[b]), n, lens[b], ptr(lat[b]),
                                                             ptr(idx[b]), C.c_void_p(_cuda_stream_ptr())))
        return lat.transpose(1, 2), idx


@torch.no_grad()
def calculate_timbre_latent(timbre_encoder: SpeakerEncoder, audio_16k_tensor: torch.Tensor, wave_lens: torch.Tensor):
    """`InferenceWrapper.calculate_timbre_latent` (evaluations/infer_arvc.py:213-223): [B, n] -> [B, 32, 128]."""
    zq, _ = timbre_encoder.tokenize_wav(audio_16k_tensor, wave_lens)
    return zq.mT.contiguous()

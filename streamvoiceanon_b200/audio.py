"""Host audio boundary: sample-rate conversion between the capture / file rate and the model's
44.1 kHz on the GPU, with `torchaudio.functional.resample` semantics (what `synth.synth_audio_44k`, the reference GUI
and -- up to the filter design -- `librosa.load(path, sr=44100)` do on the host, evaluations/infer_arvc.py:274-278).
The filter bank is the one `torchaudio.transforms.Resample` holds (built in fp64, rounded once; the GUI's resamplers);
`torchaudio.functional.resample` on an fp32 waveform builds the same bank in fp32, ~1e-6 away (oracle/prompt.py)."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .engine import Engine, ptr, _cuda_stream_ptr


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """The polyphase filter bank of torchaudio.functional.resample (method "sinc_interp_hann", default width / rolloff),
    restated with the same torch op sequence so that the values are identical: returns (kernel [new, 2*width + orig] f32,
    width, orig, new) for the gcd-reduced ratio."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1)[:, None, None] / new + idx
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.to(torch.float32)[:, 0].contiguous(), width, orig, new


class Resampler:
    """`Resampler(16000, 44100)(wave)`: wave [..., n] float32 on the host or the device -> [..., ceil(n * new / orig)] where the
    input lives (a host input is staged through the C ABI)."""

    def __init__(self, orig_freq: int, new_freq: int, device=None):
        self._engine = Engine.get(device)
        self.kernel, self.width, self.orig, self.new = sinc_resample_kernel(orig_freq, new_freq)
        self._kernel_dev = self.kernel.to(torch.device("cuda", self._engine.device))

    @torch.no_grad()
    def __call__(self, wave: torch.Tensor) -> torch.Tensor:
        if self.orig == self.new:
            return wave
        shape = wave.shape
        rows = wave.reshape(-1, shape[-1]).float().contiguous()
        n_in = rows.shape[1]
        n_out = int(math.ceil(self.new * n_in / self.orig))
        out = torch.empty(rows.shape[0], n_out, dtype=torch.float32, device=rows.device)
        for r in range(rows.shape[0]):
            _lib.check(self._engine.lib.svanon_resample(self._engine.handle, ptr(rows[r]), n_in, ptr(self._kernel_dev), self.orig,
                                                        self.new, self.width, ptr(out[r]), n_out, C.c_void_p(_cuda_stream_ptr())))
        return out.reshape(*shape[:-1], n_out)

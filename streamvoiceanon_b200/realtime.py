"""Headless twin of the reference GUI's audio path (`evaluations/real-time-gui.py`): everything
between the sound card's block callback and `InferenceWrapper.process_one_chunk`, without Tk or PortAudio (neither is in
this image; a GUI binds `RealtimeSession.audio_callback` to `sounddevice.Stream(callback=...)` unchanged).

    custom_infer(model_set, reference_wav, name, input_wav, n_frame_delay, alpha)      real-time-gui.py:32-49
        (re-)prompts when the reference file or the block size changes -- `prefill_prompt(max_prompt_frames=64)`,
        `setup_stream_caches(encode_window_frames=64, decode_window_frames=64, max_seq_frames=768, buffer_frames=32)` --
        then one `process_one_chunk`.  `model_set` is any object with the `InferenceWrapper` surface: the engine's
        (streamvoiceanon_b200.InferenceWrapper) or the reference's own class over the shims.
    RealtimeSession.start(reference_wav, ...)                                          start_vc, :1204-1287
        resampler device rate <- model rate, warm-up chunks (n_frame_delay + 3 silent blocks), state reset.
    RealtimeSession.audio_callback(indata [frames, channels], outdata, ...)            audio_callback, :1316-1359
        mono mix, block shift-in, conversion, resampling to the device rate on the GPU (svanon_resample), channel fan-out.

State that the reference keeps in module globals (`reference_wav_name`, `decode_chunk_frames`) lives in a `GuiState`;
the module-level `custom_infer` uses one shared instance, like the reference."""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

SAMPLES_PER_FRAME = 2048


class GuiState:
    def __init__(self):
        self.reference_wav_name = ""
        self.decode_chunk_frames = 0


_STATE = GuiState()


@torch.no_grad()
def custom_infer(model_set, reference_wav, new_reference_wav_name, input_wav, n_frame_delay=2, alpha=0.7,
                 state: Optional[GuiState] = None, device=None):
    """real-time-gui.py:32-49, line for line in behaviour: block [chunk * 2048] in -> converted block (CPU tensor) out."""
    st = state if state is not None else _STATE
    device = device if device is not None else getattr(model_set, "device", None)
    chunk = input_wav.size(-1) // SAMPLES_PER_FRAME
    if st.reference_wav_name != new_reference_wav_name or st.decode_chunk_frames != chunk:
        ref = reference_wav if torch.is_tensor(reference_wav) else torch.from_numpy(np.asarray(reference_wav))
        model_set.prefill_prompt(ref.to(device).unsqueeze(0), max_prompt_frames=64, delay=n_frame_delay, alpha=alpha)
        model_set.setup_stream_caches(encode_window_frames=64, decode_window_frames=64, max_seq_frames=768, buffer_frames=32,
                                      decode_chunk_frames=chunk)
        st.reference_wav_name = new_reference_wav_name
        st.decode_chunk_frames = chunk
    pred_wave = model_set.process_one_chunk(input_wav.to(device).unsqueeze(0))
    return pred_wave.squeeze().cpu()


class RealtimeSession:
    """The GUI's `start_vc` + `audio_callback` around a model set.  `samplerate` is the sound device's rate (the GUI's
    "sr_device" choice) or the model's 44.1 kHz ("sr_model"); `block_frame` the block size in 2048-sample frames."""

    def __init__(self, model_set, samplerate: Optional[int] = None, channels: int = 1, block_frame: int = 1,
                 n_frame_delay: int = 2, alpha: float = 0.7):
        self.model_set = model_set
        self.model_sr = int(getattr(model_set, "sr", 44100))
        self.samplerate = int(samplerate) if samplerate else self.model_sr
        self.channels = int(channels)
        self.block_frame = int(block_frame * SAMPLES_PER_FRAME)
        self.n_frame_delay, self.alpha = int(n_frame_delay), float(alpha)
        self.device = getattr(model_set, "device", torch.device("cuda"))
        self.state = GuiState()
        self.function = "vc"
        self.reference_wav = None
        self.reference_name = ""
        self.infer_ms = 0.0
        self.resampler2 = None

    def start(self, reference_wav, reference_name: str = "reference"):
        """start_vc (:1204-1287): buffers, the output resampler, warm-up with silent blocks, then the prompt state is
        invalidated so that the first real block prompts again (the reference resets `reference_wav_name`)."""
        from .audio import Resampler
        self.reference_wav = reference_wav if torch.is_tensor(reference_wav) else torch.from_numpy(np.asarray(reference_wav))
        self.reference_name = reference_name
        self.input_wav = torch.zeros(self.block_frame, device=self.device, dtype=torch.float32)
        self.resampler2 = Resampler(self.model_sr, self.samplerate) if self.model_sr != self.samplerate else None
        dummy = torch.zeros(self.block_frame, device=self.device, dtype=torch.float32)
        for _ in range(self.n_frame_delay + 3):
            custom_infer(self.model_set, self.reference_wav, reference_name, dummy, self.n_frame_delay, self.alpha,
                         state=self.state, device=self.device)
        self.state.reference_wav_name = ""

    def audio_callback(self, indata: np.ndarray, outdata: np.ndarray, frames=None, times=None, status=None):
        """audio_callback (:1316-1359): indata [frames, channels] float32 at the device rate (the GUI opens the stream with
        blocksize = block_frame samples), outdata [frames, channels] written in place."""
        t0 = time.perf_counter()
        mono = indata.T if indata.ndim == 2 else indata[None]
        mono = mono[0] if mono.shape[0] == 1 else mono.mean(axis=0)                  # librosa.to_mono
        self.input_wav[:-self.block_frame] = self.input_wav[self.block_frame:].clone()
        self.input_wav[-mono.shape[0]:] = torch.from_numpy(np.ascontiguousarray(mono, dtype=np.float32)).to(self.device)
        if self.function == "vc":
            infer_wav = custom_infer(self.model_set, self.reference_wav, self.reference_name, self.input_wav,
                                     self.n_frame_delay, self.alpha, state=self.state, device=self.device)
            if self.resampler2 is not None:
                infer_wav = self.resampler2(infer_wav)
        else:
            infer_wav = self.input_wav.clone()
        outdata[:] = infer_wav[: self.block_frame][None].repeat(self.channels, 1).t().cpu().numpy()
        self.infer_ms = (time.perf_counter() - t0) * 1e3

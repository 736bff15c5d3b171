"""Import shim (test infrastructure only) for `librosa.load`/`to_mono` used at
evaluations/infer_arvc.py:254,274,615,624.  WAV files only; resampling through
torchaudio.functional.resample when the file rate differs from `sr`."""
import numpy as np
from scipy.io import wavfile


def to_mono(y):
    return y if y.ndim == 1 else y.mean(axis=0)


def load(path, sr=None, mono=True):
    rate, data = wavfile.read(str(path))
    if data.dtype == np.int16:
        data = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float32) / 2147483648.0
    else:
        data = data.astype(np.float32)
    if data.ndim == 2:
        data = data.T
        if mono:
            data = data.mean(axis=0)
    if sr is not None and sr != rate:
        import torch
        import torchaudio.functional as AF
        data = AF.resample(torch.from_numpy(np.ascontiguousarray(data)), rate, sr).numpy()
        rate = sr
    return np.ascontiguousarray(data, dtype=np.float32), rate

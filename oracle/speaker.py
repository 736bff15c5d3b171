"""CPU restatement of the style-vector branch of the prompt path.  TEST INFRASTRUCTURE ONLY --
never imported by the product; groundwork for the CUDA port of this row (no CUDA counterpart yet).

    wave (16 kHz) -> kaldi fbank (80 bins) -> minus time-mean -> CAMPPlus -> style vector [192]
    (`InferenceWrapper.calculate_style_vec`, evaluations/infer_arvc.py:179-211)

Pinned: tests/golden/style_vec.npz holds features and embeddings of the UNMODIFIED reference
(`torchaudio.compliance.kaldi.fbank` + `modules.campplus.DTDNN.CAMPPlus` through the reference's own
`calculate_style_vec`), written by oracle/make_golden_style.py; tests/test_oracle_golden.py checks this file against them.

Third-party arithmetic: the filterbank lives in torchaudio (reference pin `torchaudio==2.4.0`, requirements.txt:7; 2.11.0
in the build container, same algorithm: a port of Kaldi's `compute-fbank-feats`).  `kaldi_fbank` restates it for the one
argument set the reference uses (num_mel_bins=80, dither=0, sample_frequency=16000, everything else default)."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ kaldi fbank
def _mel(f):
    return 1127.0 * math.log(1.0 + f / 700.0)


def kaldi_mel_banks(num_bins: int = 80, padded: int = 512, sr: float = 16000.0, low: float = 20.0) -> torch.Tensor:
    """`get_mel_banks` of torchaudio.compliance.kaldi (no VTLN warp): triangles that are linear in the MEL domain between
    mel(low) and mel(nyquist); [num_bins, padded/2 + 1] with the zero column torchaudio pads for the Nyquist bin."""
    n_fft_bins = padded // 2
    mel_low, mel_high = _mel(low), _mel(0.5 * sr)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = 1127.0 * (1.0 + (sr / padded) * torch.arange(n_fft_bins) / 700.0).log()
    mel = mel.unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return F.pad(bins, (0, 1), value=0.0)


def kaldi_fbank(wave: torch.Tensor, num_mel_bins: int = 80, sr: float = 16000.0) -> torch.Tensor:
    """wave [1, n] float32 -> [m, num_mel_bins], m = 1 + (n - 400) // 160 (snip_edges).  Per 25 ms frame (10 ms hop):
    remove the DC offset, pre-emphasis 0.97 (first sample against itself), povey window (hann(400, symmetric) ** 0.85),
    zero-pad to 512, power spectrum, mel filterbank, log(max(., float32 eps)).  infer_arvc.py:186-191 call site."""
    win, shift, padded = int(sr * 0.025), int(sr * 0.010), 512
    x = wave[0].to(torch.float32)
    n = x.numel()
    if n < win:
        return torch.empty(0, num_mel_bins)
    m = 1 + (n - win) // shift
    frames = x.as_strided((m, win), (shift, 1)).clone()
    frames = frames - frames.mean(dim=1, keepdim=True)
    prev = F.pad(frames.unsqueeze(0), (1, 0), mode="replicate").squeeze(0)
    frames = frames - 0.97 * prev[:, :-1]
    window = torch.hann_window(win, periodic=False).pow(0.85)
    frames = frames * window.unsqueeze(0)
    frames = F.pad(frames, (0, padded - win))
    spec = torch.fft.rfft(frames).abs().pow(2.0)
    mel = torch.mm(spec, kaldi_mel_banks(num_mel_bins, padded, sr).T)
    return torch.max(mel, torch.tensor(torch.finfo(torch.float32).eps)).log()


# ------------------------------------------------------------------------------------------------ CAMPPlus
def _bn(x, sd, key, affine=True):
    """eval-mode BatchNorm (layers.py:17-21): per-channel affine map from the running statistics, eps 1e-5."""
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd.get(key + ".weight") if affine else None,
                        sd.get(key + ".bias") if affine else None, False, 0.0, 1e-5)


def _res_block(x, sd, pre, stride):
    """BasicResBlock (layers.py:223-266): conv3x3(stride (s,1)) BN ReLU conv3x3 BN + shortcut (1x1 conv + BN when strided)."""
    out = F.relu(_bn(F.conv2d(x, sd[pre + ".conv1.weight"], None, (stride, 1), 1), sd, pre + ".bn1"))
    out = _bn(F.conv2d(out, sd[pre + ".conv2.weight"], None, 1, 1), sd, pre + ".bn2")
    if stride != 1:
        x = _bn(F.conv2d(x, sd[pre + ".shortcut.0.weight"], None, (stride, 1)), sd, pre + ".shortcut.1")
    return F.relu(out + x)


def _fcm(x, sd):
    """FCM front end (DTDNN.py:39-48): [B, 80, T] -> [B, 32 * 10, T]; frequency is strided 2 x 2 x 2, time is not."""
    out = F.relu(_bn(F.conv2d(x.unsqueeze(1), sd["head.conv1.weight"], None, 1, 1), sd, "head.bn1"))
    for layer in (1, 2):
        out = _res_block(out, sd, f"head.layer{layer}.0", 2)
        out = _res_block(out, sd, f"head.layer{layer}.1", 1)
    out = F.relu(_bn(F.conv2d(out, sd["head.conv2.weight"], None, (2, 1), 1), sd, "head.bn2"))
    return out.reshape(out.shape[0], out.shape[1] * out.shape[2], out.shape[3])


def _seg_pooling(x, seg_len=100):
    """CAMLayer.seg_pooling (layers.py:113-123): mean over segments of 100 frames (last one short), repeated back."""
    seg = F.avg_pool1d(x, kernel_size=seg_len, stride=seg_len, ceil_mode=True)
    seg = seg.unsqueeze(-1).expand(*seg.shape, seg_len).reshape(*seg.shape[:-1], -1)
    return seg[..., : x.shape[-1]]


def _cam_dense_layer(x, sd, pre, dilation):
    """CAMDenseTDNNLayer (layers.py:126-165): BN ReLU 1x1 conv -> BN ReLU -> CAM layer (local k3 dilated conv times a
    sigmoid mask computed from global mean + segment mean context, layers.py:84-111)."""
    h = F.conv1d(F.relu(_bn(x, sd, pre + ".nonlinear1.batchnorm")), sd[pre + ".linear1.weight"])
    h = F.relu(_bn(h, sd, pre + ".nonlinear2.batchnorm"))
    y = F.conv1d(h, sd[pre + ".cam_layer.linear_local.weight"], None, 1, dilation, dilation)
    ctx = h.mean(-1, keepdim=True) + _seg_pooling(h)
    ctx = F.relu(F.conv1d(ctx, sd[pre + ".cam_layer.linear1.weight"], sd[pre + ".cam_layer.linear1.bias"]))
    mask = torch.sigmoid(F.conv1d(ctx, sd[pre + ".cam_layer.linear2.weight"], sd[pre + ".cam_layer.linear2.bias"]))
    return y * mask


def campplus_forward(feat: torch.Tensor, lens: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """CAMPPlus.forward (DTDNN.py:132-138): feat [B, T, 80], lens [B] (valid frames AFTER the stride-2 TDNN) -> [B, 192]."""
    x = _fcm(feat.permute(0, 2, 1), sd)
    x = F.relu(_bn(F.conv1d(x, sd["xvector.tdnn.linear.weight"], None, 2, 2), sd, "xvector.tdnn.nonlinear.batchnorm"))
    for i, (num_layers, dilation) in enumerate(zip((12, 24, 16), (1, 2, 2))):
        for j in range(num_layers):
            x = torch.cat([x, _cam_dense_layer(x, sd, f"xvector.block{i + 1}.tdnnd{j + 1}", dilation)], dim=1)
        x = F.conv1d(F.relu(_bn(x, sd, f"xvector.transit{i + 1}.nonlinear.batchnorm")), sd[f"xvector.transit{i + 1}.linear.weight"])
    x = F.relu(_bn(x, sd, "xvector.out_nonlinear.batchnorm"))
    stats = []                                                   # masked_statistics_pooling, layers.py:34-44
    for i in range(x.shape[0]):
        xi = x[i, :, : int(lens[i])]
        stats.append(torch.cat([xi.mean(dim=-1), xi.std(dim=-1, unbiased=True)], dim=-1))
    x = torch.stack(stats, dim=0)
    x = F.conv1d(x.unsqueeze(-1), sd["dense.linear.weight"]).squeeze(-1)
    return _bn(x, sd, "dense.nonlinear.batchnorm", affine=False)


def calculate_style_vec(wave16k: torch.Tensor, wave_lens: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """infer_arvc.py:179-211: per row fbank of the valid samples minus its time-mean; rows padded to the longest with
    the row's minimum; lens = frames // 2; CAMPPlus."""
    feats = []
    for b in range(wave16k.shape[0]):
        f = kaldi_fbank(wave16k[b: b + 1, : int(wave_lens[b])])
        feats.append(f - f.mean(dim=0, keepdim=True))
    longest = max(f.shape[0] for f in feats)
    lens = torch.tensor([f.shape[0] for f in feats], dtype=torch.int32) // 2
    feats = [F.pad(f, (0, 0, 0, longest - f.shape[0]), value=float(f.min())) for f in feats]
    return campplus_forward(torch.stack(feats, dim=0), lens, sd)

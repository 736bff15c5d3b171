"""CPU restatement of the two speaker-embedding branches of the prompt path.  TEST
INFRASTRUCTURE ONLY -- never imported by the product; groundwork for the CUDA port of this row (no CUDA counterpart yet).

    wave (16 kHz) -> kaldi fbank (80 bins) -> minus time-mean -> CAMPPlus -> style vector [192]
    (`InferenceWrapper.calculate_style_vec`, evaluations/infer_arvc.py:179-211)
    wave (16 kHz) -> slaney mel magnitudes (128 bins, 20 ms hop) -> ECAPA-TDNN trunk -> PerceiverResampler (32 latents)
    -> FSQ (levels 4^6) -> timbre latents [32, 128]
    (`InferenceWrapper.calculate_timbre_latent`, evaluations/infer_arvc.py:213-223; `SpeakerEncoder.tokenize_wav`,
    modules/bicodec_speaker_encoder/speaker_encoder.py:136-144)

Pinned: tests/golden/style_vec.npz and tests/golden/timbre_latent.npz hold features and embeddings of the UNMODIFIED
reference (`torchaudio.compliance.kaldi.fbank` + `modules.campplus.DTDNN.CAMPPlus` through the reference's own
`calculate_style_vec`; `torchaudio.transforms.MelSpectrogram` + `SpeakerEncoder.tokenize_wav` through
`calculate_timbre_latent`), written by oracle/make_golden_style.py; tests/test_oracle_golden.py checks this file against
them.

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


# ------------------------------------------------------------------------------------------------ timbre branch
def _hz_to_mel_slaney(f: float) -> float:
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    if f >= min_log_hz:
        return min_log_hz / f_sp + math.log(f / min_log_hz) / (math.log(6.4) / 27.0)
    return f / f_sp


def slaney_mel_fbanks(n_freqs: int = 513, f_min: float = 10.0, f_max: float = 8000.0, n_mels: int = 128,
                      sr: int = 16000) -> torch.Tensor:
    """`torchaudio.functional.melscale_fbanks(..., norm="slaney", mel_scale="slaney")`: triangles that are linear in Hz
    between slaney-mel-spaced points, each scaled by 2 / (its band width)  ->  [n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sr // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
    f_pts = f_sp * m_pts
    log_t = m_pts >= min_log_mel
    f_pts[log_t] = min_log_hz * torch.exp(logstep * (m_pts[log_t] - min_log_mel))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    return fb * (2.0 / (f_pts[2: n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)


def timbre_mel(wave16k: torch.Tensor) -> torch.Tensor:
    """The `mel_fn` of configs/hydra_arcs/sv/sparktts_speaker_encoder.yaml (torchaudio MelSpectrogram: n_fft 1024, hann
    window 640 (periodic) centred in the frame, hop 320, reflect-padded centre frames, MAGNITUDE (power 1), slaney mel
    128 bins from 10 Hz): wave [B, n] -> [B, T, 128], T = n // 320 + 1  (speaker_encoder.py:138)."""
    spec = torch.stft(wave16k, 1024, hop_length=320, win_length=640, window=torch.hann_window(640), center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True).abs()
    return torch.matmul(spec.transpose(-1, -2), slaney_mel_fbanks())


def _conv_relu_bn(x, sd, pre, padding=0, dilation=1):
    """Conv1dReluBn (ecapa_tdnn.py:69-90): conv -> ReLU -> BatchNorm, in this order."""
    return _bn(F.relu(F.conv1d(x, sd[pre + ".conv.weight"], sd[pre + ".conv.bias"], 1, padding, dilation)), sd, pre + ".bn")


def _se_res2block(x, sd, pre, dilation):
    """SE_Res2Block (ecapa_tdnn.py:117-133): 1x1 -> Res2 (8 splits of 64, 7 dilated k3 convs chained, ecapa_tdnn.py:12-63)
    -> 1x1 -> squeeze-excitation gate (ecapa_tdnn.py:96-111), plus the residual."""
    h = _conv_relu_bn(x, sd, pre + ".0")
    spx = torch.split(h, 64, 1)
    out, sp = [], spx[0]
    for i in range(7):
        if i >= 1:
            sp = sp + spx[i]
        sp = F.conv1d(sp, sd[f"{pre}.1.convs.{i}.weight"], sd[f"{pre}.1.convs.{i}.bias"], 1, dilation, dilation)
        sp = _bn(F.relu(sp), sd, f"{pre}.1.bns.{i}")
        out.append(sp)
    out.append(spx[7])
    h = _conv_relu_bn(torch.cat(out, dim=1), sd, pre + ".2")
    g = F.relu(F.linear(h.mean(dim=2), sd[pre + ".3.linear1.weight"], sd[pre + ".3.linear1.bias"]))
    g = torch.sigmoid(F.linear(g, sd[pre + ".3.linear2.weight"], sd[pre + ".3.linear2.bias"]))
    return x + h * g.unsqueeze(2)


def ecapa_latent(mel: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """ECAPA_TDNN.forward(..., return_latent=True)[1] (ecapa_tdnn.py:191-209): mel [B, T, 128] -> [B, 1536, T]."""
    e = "speaker_encoder"
    out1 = _conv_relu_bn(mel.permute(0, 2, 1), sd, e + ".layer1", padding=2)
    out2 = _se_res2block(out1, sd, e + ".layer2.se_res2block", 2)
    out3 = _se_res2block(out2, sd, e + ".layer3.se_res2block", 3)
    out4 = _se_res2block(out3, sd, e + ".layer4.se_res2block", 4)
    return F.relu(F.conv1d(torch.cat([out2, out3, out4], dim=1), sd[e + ".conv.weight"], sd[e + ".conv.bias"]))


def perceiver_resample(x: torch.Tensor, mask: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """PerceiverResampler.forward (perceiver_encoder.py:339-351): x [B, T, 1536], mask [B, 32 + T] (True = attend) ->
    [B, 32, 128].  Two layers of cross attention whose keys/values are [latents ; context] (`cross_attn_include_queries`,
    perceiver_encoder.py:275-289; 8 heads x 64, no bias, masked scores set to -float32 max, :140-155) and a GEGLU MLP
    (128 -> 2 x 341 -> 128, :207-229), both residual; final RMSNorm = normalize * sqrt(128) * gamma (:177-190)."""
    p = "perceiver_sampler"
    B = x.shape[0]
    ctx = F.linear(x, sd[p + ".proj_context.weight"], sd[p + ".proj_context.bias"])
    lat = sd[p + ".latents"].unsqueeze(0).expand(B, -1, -1)
    neg = -torch.finfo(torch.float32).max
    for layer in range(2):
        a = f"{p}.layers.{layer}.0"
        kv_in = torch.cat((lat, ctx), dim=-2)
        q = F.linear(lat, sd[a + ".to_q.weight"])
        k, v = F.linear(kv_in, sd[a + ".to_kv.weight"]).chunk(2, dim=-1)
        q, k, v = (t.reshape(B, t.shape[1], 8, 64).permute(0, 2, 1, 3) for t in (q, k, v))
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * (64 ** -0.5)
        sim = sim.masked_fill(~mask[:, None, None, :], neg)
        out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
        out = out.permute(0, 2, 1, 3).reshape(B, -1, 512)
        lat = F.linear(out, sd[a + ".to_out.weight"]) + lat
        f = f"{p}.layers.{layer}.1"
        h, gate = F.linear(lat, sd[f + ".0.weight"], sd[f + ".0.bias"]).chunk(2, dim=-1)
        lat = F.linear(F.gelu(gate) * h, sd[f + ".2.weight"], sd[f + ".2.bias"]) + lat
    return F.normalize(lat, dim=-1) * (128 ** 0.5) * sd[p + ".norm.gamma"]


def fsq4_quantize(z: torch.Tensor):
    """FSQ with levels [4] * 6 (fsq/finite_scalar_quantization.py:126-162): bound = tanh(z + atanh(0.5 / h)) * h - 0.5 with
    h = 1.5 * 1.001, round, / 2  ->  codes in {-1, -0.5, 0, 0.5}; index = sum((code * 2 + 2) * 4^i).  z [..., 6]."""
    half_l = torch.full((6,), 3.0) * (1 + 1e-3) / 2
    offset = torch.full((6,), 0.5)
    shift = (offset / half_l).atanh()
    bounded = (z + shift).tanh() * half_l - offset
    codes = bounded.round() / 2
    basis = torch.tensor([1, 4, 16, 64, 256, 1024], dtype=torch.int32)
    indices = ((codes * 2 + 2) * basis).sum(dim=-1).to(torch.int32)
    return codes, indices, bounded


def calculate_timbre_latent(wave16k: torch.Tensor, wave_lens: torch.Tensor, sd: Dict[str, torch.Tensor]):
    """infer_arvc.py:213-223 -> `tokenize_wav` (speaker_encoder.py:136-144): returns (timbre latents [B, 32, 128] =
    `zq.mT`, FSQ indices [B, 32] int32, pre-rounding FSQ coordinates [B, 32, 6]).  The attention mask keeps the 32
    latent slots and the first wave_len // 320 context frames."""
    mel = timbre_mel(wave16k)
    feats = ecapa_latent(mel, sd)
    T = feats.shape[2]
    mel_lens = wave_lens // 320
    mask = torch.arange(T + 32).unsqueeze(0) < (mel_lens + 32).unsqueeze(1)
    lat = perceiver_resample(feats.transpose(1, 2), mask, sd)
    z = F.linear(lat, sd["quantizer.project_in.weight"], sd["quantizer.project_in.bias"])
    codes, indices, bounded = fsq4_quantize(z)
    zq = F.linear(codes, sd["quantizer.project_out.weight"], sd["quantizer.project_out.bias"])
    return zq, indices, bounded

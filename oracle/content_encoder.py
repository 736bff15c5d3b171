"""Stage E oracle: waveform -> 13-bit BSQ content ids.  TEST INFRASTRUCTURE ONLY.

Restates, in plain torch fp32 on CPU, `FireflyArchitecture.encode`
(modules/vqgan/modules/firefly_encoder.py:553-566) for the tokenizer config
configs/hydra_arcs/speech_tokenizers/causal-encoder-lfq-8192.yaml.  Weights are passed
as a flat state-dict with the reference's key names.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

N_FFT = 2048
HOP = 512
N_MELS = 160
SR = 44100
DIMS = [128, 256, 384, 512]
DEPTHS = [3, 3, 9, 3]


def slaney_fbanks(n_freqs=N_FFT // 2 + 1, f_min=0.0, f_max=SR // 2, n_mels=N_MELS, sample_rate=SR):
    """torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney") as used at
    modules/vqgan/spectrogram.py:93-106, restated (returns [n_freqs, n_mels])."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)

    def hz_to_mel(f):
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = math.log(6.4) / 27.0
        if f >= min_log_hz:
            mels = min_log_mel + math.log(f / min_log_hz) / logstep
        return mels

    def mel_to_hz(mels):
        f_sp = 200.0 / 3
        freqs = f_sp * mels
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = math.log(6.4) / 27.0
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
        return freqs

    m_min, m_max = hz_to_mel(f_min), hz_to_mel(float(f_max))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = mel_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2 : n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


_FB = None


def log_mel(wav: torch.Tensor) -> torch.Tensor:
    """modules/vqgan/spectrogram.py:26-65 (LinearSpectrogram) and :108-130
    (LogMelSpectrogram): left-pad win-hop zeros, STFT(center=False, periodic Hann),
    sqrt(re^2+im^2+1e-6), slaney filterbank, log(clamp(.,1e-5)).  [B,L] -> [B,160,L/512]."""
    global _FB
    if _FB is None:
        _FB = slaney_fbanks()
    y = F.pad(wav.float(), (N_FFT - HOP, 0))
    spec = torch.stft(y, N_FFT, hop_length=HOP, win_length=N_FFT, window=torch.hann_window(N_FFT),
                      center=False, normalized=False, onesided=True, return_complex=True)
    spec = torch.view_as_real(spec)
    lin = torch.sqrt(spec.pow(2).sum(-1) + 1e-6)
    mel = torch.matmul(lin.transpose(-1, -2), _FB).transpose(-1, -2)
    return torch.log(torch.clamp(mel, min=1e-5))


def causal_conv1d(x, w, b, stride=1, dilation=1, groups=1):
    """FishConvNet.forward, modules/vqgan/modules/firefly.py:92-103: left pad
    (k-1)*d+1-stride zeros, then a plain Conv1d."""
    k_eff = (w.shape[-1] - 1) * dilation + 1
    x = F.pad(x, (k_eff - stride, 0))
    return F.conv1d(x, w, b, stride=stride, dilation=dilation, groups=groups)


def layer_norm_cf(x, w, b, eps=1e-6):
    """channels_first LayerNorm, modules/vqgan/modules/firefly.py:366-371."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None] * x + b[:, None]


def convnext_block(x, sd, p):
    """ConvNeXtBlock.forward, modules/vqgan/modules/firefly.py:421-440."""
    dim = x.shape[1]
    h = causal_conv1d(x, sd[p + ".dwconv.conv.weight"], sd[p + ".dwconv.conv.bias"], groups=dim)
    h = h.permute(0, 2, 1)
    h = F.layer_norm(h, (dim,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    h = F.linear(h, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[p + ".pwconv2.weight"], sd[p + ".pwconv2.bias"])
    h = sd[p + ".gamma"] * h
    return x + h.permute(0, 2, 1)


def convnext_encoder(x, sd, p="backbone"):
    """ConvNeXtEncoder.forward, modules/vqgan/modules/firefly.py:506-517 (gin_channels=0)."""
    for i in range(4):
        if i == 0:
            x = causal_conv1d(x, sd[f"{p}.downsample_layers.0.0.conv.weight"], sd[f"{p}.downsample_layers.0.0.conv.bias"])
            x = layer_norm_cf(x, sd[f"{p}.downsample_layers.0.1.weight"], sd[f"{p}.downsample_layers.0.1.bias"])
        else:
            x = layer_norm_cf(x, sd[f"{p}.downsample_layers.{i}.0.weight"], sd[f"{p}.downsample_layers.{i}.0.bias"])
            x = F.conv1d(x, sd[f"{p}.downsample_layers.{i}.1.weight"], sd[f"{p}.downsample_layers.{i}.1.bias"])
        for j in range(DEPTHS[i]):
            x = convnext_block(x, sd, f"{p}.stages.{i}.{j}")
    return layer_norm_cf(x, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"])


def rope_table(seq_len, n_elem=64, base=10000.0):
    """precompute_freqs_cis, modules/vqgan/windowed_transformer.py:356-365: the cos/sin
    table is rounded to bfloat16 and later multiplied in fp32."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: n_elem // 2].float() / n_elem))
    t = torch.arange(seq_len)
    freqs = torch.outer(t, freqs)
    cis = torch.polar(torch.ones_like(freqs), freqs)
    return torch.stack([cis.real, cis.imag], dim=-1).to(torch.bfloat16)


def apply_rope(x, table):
    """apply_rotary_emb, windowed_transformer.py:368-380.  x [B,S,H,D], table [S,D/2,2]."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    fc = table.view(1, xs.size(1), 1, xs.size(3), 2)
    out = torch.stack([xs[..., 0] * fc[..., 0] - xs[..., 1] * fc[..., 1],
                       xs[..., 1] * fc[..., 0] + xs[..., 0] * fc[..., 1]], -1)
    return out.flatten(3).type_as(x)


def rms_norm(x, w, eps=1e-5):
    """RMSNorm.forward, windowed_transformer.py:248-259."""
    return (x.float() * torch.rsqrt(torch.mean(x.float() * x.float(), dim=-1, keepdim=True) + eps)).type_as(x) * w


def window_mask(n, window=512):
    """make_window_limited_mask (causal branch), windowed_transformer.py:291-303."""
    row = torch.arange(n).view(-1, 1)
    col = torch.arange(n)
    return (col >= (row - window + 1).clamp(min=0)) & (col <= row)


def window_transformer(x, sd, p="quantizer.pre_module", n_layer=8, n_head=8, window=512):
    """WindowLimitedTransformer.forward, windowed_transformer.py:337-354 -> Transformer.forward
    :103-120 -> TransformerBlock :134-143 -> Attention :163-194, FeedForward :244-245.
    x is channels-first [B,C,T]."""
    x = x.transpose(1, 2)
    B, S, D = x.shape
    hd = D // n_head
    mask = window_mask(S, window)[None, None]
    table = rope_table(2048)[:S]
    for i in range(n_layer):
        lp = f"{p}.layers.{i}"
        h = rms_norm(x, sd[lp + ".attention_norm.weight"])
        q, k, v = F.linear(h, sd[lp + ".attention.wqkv.weight"]).split([D, D, D], dim=-1)
        q = apply_rope(q.view(B, S, n_head, hd), table).transpose(1, 2)
        k = apply_rope(k.view(B, S, n_head, hd), table).transpose(1, 2)
        v = v.view(B, S, n_head, hd).transpose(1, 2)
        y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        y = y.transpose(1, 2).contiguous().view(B, S, D)
        y = F.linear(y, sd[lp + ".attention.wo.weight"])
        x = x + y * sd[lp + ".attention_layer_scale.gamma"]
        h = rms_norm(x, sd[lp + ".ffn_norm.weight"])
        f = F.linear(F.silu(F.linear(h, sd[lp + ".feed_forward.w1.weight"])) * F.linear(h, sd[lp + ".feed_forward.w3.weight"]),
                     sd[lp + ".feed_forward.w2.weight"])
        x = x + f * sd[lp + ".ffn_layer_scale.gamma"]
    x = rms_norm(x, sd[p + ".norm.weight"])
    return x.transpose(1, 2)


def bsq_ids(z, sd, p="quantizer.residual_bsq.rvqs.0"):
    """LFQ.forward inference arithmetic, modules/vqgan/modules/bsq.py:330-369 (one group,
    one codebook of 13 bits): project_in, l2-normalise, bit = x > 0,
    id = sum(bit_i * 2^(12-i)).  z [B,T,512] -> int64 [B,T]."""
    x = F.linear(z, sd[p + ".project_in.weight"], sd[p + ".project_in.bias"])
    x = F.normalize(x, dim=-1)
    bits = (x > 0).long()
    weights = 2 ** torch.arange(12, -1, -1)
    return (bits * weights).sum(-1)


def encode_features(wav, sd):
    """Everything of `encode()` up to the BSQ projection input (for error budgets)."""
    mels = log_mel(wav)
    x = convnext_encoder(mels, sd)
    for i in range(2):
        # DownsampleBinarySphericalQuantize.downsample, bsq_no_upsample.py:46-60
        x = causal_conv1d(x, sd[f"quantizer.downsample.{i}.0.conv.weight"], sd[f"quantizer.downsample.{i}.0.conv.bias"], stride=2)
        x = convnext_block(x, sd, f"quantizer.downsample.{i}.1")
    return window_transformer(x, sd)


def encode(wav, sd, lens=None):
    """FireflyArchitecture.encode (firefly_encoder.py:553-566) +
    DownsampleBinarySphericalQuantize.encode (bsq_no_upsample.py:103-107).
    wav [B,L] -> (ids int64 [1,B,T], feature_lengths).  The mel mask is all-ones when
    lens == L (the streaming case), and is applied exactly like the reference otherwise."""
    B, L = wav.shape
    if lens is None:
        lens = torch.full((B,), L, dtype=torch.long)
    mels = log_mel(wav)
    mel_lens = lens // HOP
    mask = (torch.arange(mels.shape[2])[None] < mel_lens[:, None])[:, None, :].float()
    mels = mels * mask
    x = convnext_encoder(mels, sd) * mask
    for i in range(2):
        x = causal_conv1d(x, sd[f"quantizer.downsample.{i}.0.conv.weight"], sd[f"quantizer.downsample.{i}.0.conv.bias"], stride=2)
        x = convnext_block(x, sd, f"quantizer.downsample.{i}.1")
    z = window_transformer(x, sd)
    ids = bsq_ids(z.transpose(1, 2), sd)
    return ids[None], mel_lens // 4

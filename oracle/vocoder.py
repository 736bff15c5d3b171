"""Stage V oracle: 8 FSQ codec ids per frame -> waveform.  TEST INFRASTRUCTURE ONLY.

Plain torch-fp32 CPU restatement of `code2wav_fn` (evaluations/infer_arvc.py:173-176):
`firefly.head(firefly.quantizer.decode(codes))` for
configs/hydra_arcs/vocoders/firefly_gan_vq.yaml.  Takes the FOLDED state-dict (weight
norm removed, as after `remove_parametrizations()`, infer_arvc.py:94).

The FSQ index arithmetic lives in the third-party `vector-quantize-pytorch==1.14.24`
(reference requirements.txt:26, call site modules/vqgan/modules/fsq.py:112-116), absent
here; it is restated from the reference's vendored twin
modules/bicodec_speaker_encoder/fsq/finite_scalar_quantization.py:143-162 and
residual_fsq.py:112-156, and pinned by tests/golden (generated through that twin).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .content_encoder import causal_conv1d, convnext_block

LEVELS = torch.tensor([8, 5, 5, 5])
BASIS = torch.tensor([1, 8, 40, 200])
UP_RATES = [8, 8, 2, 2, 2]
UP_KERNELS = [16, 16, 4, 4, 4]
RES_KERNELS = [3, 7, 11]
RES_DILATIONS = [1, 3, 5]


def fsq_decode(indices, sd):
    """GroupedResidualFSQ.get_output_from_indices for 8 groups x 1 quantizer, levels
    (8,5,5,5): digit_i = (idx // basis_i) % level_i; code = (digit - L//2) / (L//2);
    scale (levels-1)^-0 = 1; project_out_g Linear 4->64; concat groups.
    indices [B,8,T] -> [B,T,512]."""
    outs = []
    for g in range(8):
        idx = indices[:, g].long().unsqueeze(-1)
        digits = (idx // BASIS) % LEVELS
        half = LEVELS // 2
        codes = (digits - half).float() / half.float()
        outs.append(F.linear(codes, sd[f"quantizer.residual_fsq.rvqs.{g}.project_out.weight"],
                             sd[f"quantizer.residual_fsq.rvqs.{g}.project_out.bias"]))
    return torch.cat(outs, dim=-1)


def trans_conv(x, w, b, k, stride):
    """FishTransConvNet.forward, modules/vqgan/modules/firefly.py:114-138."""
    if stride == k // 2:
        x = F.pad(x, (1, 0))
    elif stride == k:
        x = F.pad(x, (1, 1))
    return F.conv_transpose1d(x, w, b, stride=stride, padding=stride, output_padding=stride % 2)


def quantizer_decode(indices, sd):
    """DownsampleFiniteScalarQuantize.decode, modules/vqgan/modules/fsq.py:112-116 and
    the upsample stack :61-74.  [B,8,T] -> [B,512,4T]."""
    z = fsq_decode(indices, sd).transpose(1, 2)
    for i in range(2):
        z = trans_conv(z, sd[f"quantizer.upsample.{i}.0.conv.weight"], sd[f"quantizer.upsample.{i}.0.conv.bias"], 2, 2)
        z = convnext_block(z, sd, f"quantizer.upsample.{i}.1")
    return z


def res_block(x, sd, p, k):
    """ResBlock1.forward, modules/vqgan/modules/firefly.py:183-190."""
    for j, d in enumerate(RES_DILATIONS):
        xt = F.silu(x)
        xt = causal_conv1d(xt, sd[f"{p}.convs1.{j}.conv.weight"], sd[f"{p}.convs1.{j}.conv.bias"], dilation=d)
        xt = F.silu(xt)
        xt = causal_conv1d(xt, sd[f"{p}.convs2.{j}.conv.weight"], sd[f"{p}.convs2.{j}.conv.bias"], dilation=d)
        x = xt + x
    return x


def head(z, sd):
    """HiFiGANGenerator.forward, modules/vqgan/modules/firefly.py:280-293.
    [B,512,L] -> [B,1,512 L]."""
    x = causal_conv1d(z, sd["head.conv_pre.conv.weight"], sd["head.conv_pre.conv.bias"])
    for i in range(5):
        x = F.silu(x)
        x = trans_conv(x, sd[f"head.ups.{i}.conv.weight"], sd[f"head.ups.{i}.conv.bias"], UP_KERNELS[i], UP_RATES[i])
        # ParallelBlock.forward :214-215 -- mean of the three ResBlock1 outputs
        x = torch.stack([res_block(x, sd, f"head.resblocks.{i}.blocks.{j}", k) for j, k in enumerate(RES_KERNELS)], dim=0).mean(dim=0)
    x = F.silu(x)
    x = causal_conv1d(x, sd["head.conv_post.conv.weight"], sd["head.conv_post.conv.bias"])
    return torch.tanh(x)


def code2wav(codes, sd):
    """evaluations/infer_arvc.py:173-176.  codes [B,8,T] -> wave [B,1,2048 T]."""
    return head(quantizer_decode(codes, sd), sd)


# ------------------------------------------------------------------------------------------ encode path (prompt)
def fsq_encode(z, sd):
    """GroupedResidualFSQ.forward -> indices for 8 groups x 1 quantizer (vector-quantize-pytorch==1.14.24, restated from
    the reference's vendored twin modules/bicodec_speaker_encoder/fsq/finite_scalar_quantization.py:126-156 and
    residual_fsq.py:160-230): per group project_in Linear 64->4, bound (tanh with the even-level half-step shift,
    eps 1e-3), round half to even, digit = round + L//2, index = sum digit * basis.  z [B,T,512] -> int32 [B,8,T]."""
    levels = LEVELS.to(torch.int32)
    half_l = (levels - 1) * (1 + 1e-3) / 2
    offset = torch.where(levels % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    half_width = levels // 2
    out = []
    for g in range(8):
        x = F.linear(z[..., g * 64:(g + 1) * 64], sd[f"quantizer.residual_fsq.rvqs.{g}.project_in.weight"],
                     sd[f"quantizer.residual_fsq.rvqs.{g}.project_in.bias"])
        bounded = (x + shift).tanh() * half_l - offset
        zhat = bounded.round() / half_width
        digits = zhat * half_width + half_width
        out.append((digits * BASIS.to(torch.int32)).sum(dim=-1).to(torch.int32))
    return torch.stack(out, dim=1)


def wav2codes(wav, sd):
    """`wav2target_fn` (evaluations/infer_arvc.py:168-171) = FireflyArchitecture.encode (firefly.py:561-574) for
    full-length rows: log-mel -> ConvNeXt backbone -> quantizer.downsample (2 x [causal conv k2 s2 + ConvNeXt]) -> FSQ
    indices.  wav [B,L] -> int32 [B,8,L//2048]."""
    from . import content_encoder as E
    x = E.convnext_encoder(E.log_mel(wav), sd, "backbone")
    for i in range(2):
        x = causal_conv1d(x, sd[f"quantizer.downsample.{i}.0.conv.weight"], sd[f"quantizer.downsample.{i}.0.conv.bias"], stride=2)
        x = convnext_block(x, sd, f"quantizer.downsample.{i}.1")
    return fsq_encode(x.transpose(1, 2), sd)

"""Deterministic synthetic checkpoints and audio.

No pretrained checkpoint exists offline (the reference downloads them from the HF hub,
README.md:61-64), so tests, `smoke()` and `bench.py` use seeded random state-dicts that
carry exactly the reference's on-disk key names and shapes:

  AR         pretrained_checkpoints/dual_ar_delay_0_8.pth                      (flat, fp32)
  tokenizer  pretrained_checkpoints/asr_s2s_bsq_8192_causal_down_whisper.pth
  vocoder    pretrained_checkpoints/firefly-gan-vq-fsq-8x1024-21hz-generator.pth
             (head convs in weight-norm form: parametrizations.weight.original0/1)

Every tensor is drawn from its own generator seeded by crc32(key) ^ seed, so any subset
can be regenerated independently and the result does not depend on generation order.
Scales are chosen so that every branch matters numerically (O(1) layer-scale gammas,
non-trivial biases, peaked output heads) -- the reference's own init leaves layer-scale
at 1e-6/1e-2, which would hide errors in whole sub-paths.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import numpy as np
import torch

SAMPLES_PER_FRAME = 2048          # evaluations/infer_arvc.py:28
MODEL_SR = 44100                  # configs/config_firefly_arvcasr_8192_delay0_8.yaml:13
FRAME_RATE = MODEL_SR / SAMPLES_PER_FRAME


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _randn(key, seed, *shape):
    return torch.randn(*shape, generator=_gen(key, seed), dtype=torch.float32)


class _Builder:
    def __init__(self, seed: int):
        self.seed = seed
        self.sd: Dict[str, torch.Tensor] = {}

    def linear(self, key, out_f, in_f, gain=1.0, bias=False, extra=()):
        fan_in = in_f * int(np.prod(extra)) if extra else in_f
        self.sd[key + ".weight"] = _randn(key + ".weight", self.seed, out_f, in_f, *extra) * (gain / math.sqrt(fan_in))
        if bias:
            self.sd[key + ".bias"] = _randn(key + ".bias", self.seed, out_f) * 0.05

    def norm(self, key, dim, bias=False):
        self.sd[key + ".weight"] = 1.0 + 0.1 * _randn(key + ".weight", self.seed, dim)
        if bias:
            self.sd[key + ".bias"] = 0.05 * _randn(key + ".bias", self.seed, dim)

    def gamma(self, key, dim, scale):
        self.sd[key] = scale * (1.0 + 0.3 * _randn(key, self.seed, dim))

    def table(self, key, rows, dim, std):
        self.sd[key] = _randn(key, self.seed, rows, dim) * std

    def convnext(self, prefix, dim, gamma_scale=0.3):
        # modules/vqgan/modules/firefly.py:375-440
        self.gamma(prefix + ".gamma", dim, gamma_scale)
        self.sd[prefix + ".dwconv.conv.weight"] = _randn(prefix + ".dwconv.conv.weight", self.seed, dim, 1, 7) / math.sqrt(7)
        self.sd[prefix + ".dwconv.conv.bias"] = 0.05 * _randn(prefix + ".dwconv.conv.bias", self.seed, dim)
        self.norm(prefix + ".norm", dim, bias=True)
        self.linear(prefix + ".pwconv1", 4 * dim, dim, bias=True)
        self.linear(prefix + ".pwconv2", dim, 4 * dim, bias=True)

    def wn_conv(self, prefix, shape, norm_dim0, gain=1.0):
        """weight-norm form (torch.nn.utils.parametrizations.weight_norm, dim=0):
        original0 = g [shape[0],1,1], original1 = v; w = g * v / ||v||_(1,2)."""
        fan_in = norm_dim0
        v = _randn(prefix + ".parametrizations.weight.original1", self.seed, *shape) * (gain / math.sqrt(fan_in))
        g = v.flatten(1).norm(dim=1).view(-1, 1, 1) * (1.0 + 0.1 * _randn(prefix + ".parametrizations.weight.original0", self.seed, shape[0], 1, 1))
        self.sd[prefix + ".parametrizations.weight.original0"] = g
        self.sd[prefix + ".parametrizations.weight.original1"] = v

    def bias(self, key, dim, std=0.05):
        self.sd[key] = std * _randn(key, self.seed, dim)


def _transformer_layer(b: _Builder, prefix, dim, inter, layer_scale=None, resid_gain=0.5):
    b.linear(prefix + ".attention.wqkv", 3 * dim, dim, gain=1.5)
    b.linear(prefix + ".attention.wo", dim, dim, gain=resid_gain)
    b.linear(prefix + ".feed_forward.w1", inter, dim)
    b.linear(prefix + ".feed_forward.w3", inter, dim)
    b.linear(prefix + ".feed_forward.w2", dim, inter, gain=resid_gain)
    b.norm(prefix + ".ffn_norm", dim)
    b.norm(prefix + ".attention_norm", dim)
    if layer_scale is not None:
        b.gamma(prefix + ".attention_layer_scale.gamma", dim, layer_scale)
        b.gamma(prefix + ".ffn_layer_scale.gamma", dim, layer_scale)


def make_ar_state_dict(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """ARVCWrapper checkpoint; key names per modules/arvc_wrapper.py:7-23 and
    modules/dual_ar_stream.py:167-205,411-457,605-625."""
    b = _Builder(seed)
    dim, inter = 768, 2304
    b.table("embedding.weight", 8192, dim, 0.7)
    b.table("decoder.model.codebook_embeddings.weight", 8000, dim, 0.35)
    for i in range(12):
        _transformer_layer(b, f"decoder.model.layers.{i}", dim, inter)
    b.norm("decoder.model.norm", dim)
    b.linear("decoder.model.output", 8192, dim, gain=3.0)
    b.table("decoder.model.fast_embeddings.weight", 1000, dim, 0.7)
    for i in range(4):
        _transformer_layer(b, f"decoder.model.fast_layers.{i}", dim, inter)
    b.norm("decoder.model.fast_norm", dim)
    b.linear("decoder.model.fast_output", 1000, dim, gain=3.0)
    b.table("decoder.wait4start_embedding.weight", 8, dim, 0.7)
    b.table("decoder.wait4end_embedding.weight", 8, dim, 0.7)
    b.linear("context_in", dim, 128, bias=True)
    b.linear("style_in", dim, 192, bias=True)
    return b.sd


def make_tokenizer_state_dict(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Content tokenizer (only the tensors `encode()` touches); key names per
    modules/vqgan/modules/firefly.py:443-517, bsq_no_upsample.py:20-81,
    windowed_transformer.py:68-143, bsq.py:173-176."""
    b = _Builder(seed)
    dims, depths = [128, 256, 384, 512], [3, 3, 9, 3]
    b.linear("backbone.downsample_layers.0.0.conv", dims[0], 160, bias=True, extra=(7,))
    b.norm("backbone.downsample_layers.0.1", dims[0], bias=True)
    for i in range(1, 4):
        b.norm(f"backbone.downsample_layers.{i}.0", dims[i - 1], bias=True)
        b.linear(f"backbone.downsample_layers.{i}.1", dims[i], dims[i - 1], bias=True, extra=(1,))
    for s in range(4):
        for j in range(depths[s]):
            b.convnext(f"backbone.stages.{s}.{j}", dims[s])
    b.norm("backbone.norm", 512, bias=True)
    for i in range(2):
        b.linear(f"quantizer.downsample.{i}.0.conv", 512, 512, bias=True, extra=(2,))
        b.convnext(f"quantizer.downsample.{i}.1", 512)
    for i in range(8):
        _transformer_layer(b, f"quantizer.pre_module.layers.{i}", 512, 1536, layer_scale=0.5, resid_gain=1.0)
    b.norm("quantizer.pre_module.norm", 512)
    b.linear("quantizer.residual_bsq.rvqs.0.project_in", 13, 512, bias=True)
    return b.sd


def make_vocoder_state_dict(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Vocoder decode path (quantizer.decode + head), head convs in weight-norm form;
    key names per modules/vqgan/modules/firefly.py:222-301 and fsq.py:19-74."""
    b = _Builder(seed)
    for g in range(8):
        b.linear(f"quantizer.residual_fsq.rvqs.{g}.project_out", 64, 4, gain=2.0, bias=True)
    for i in range(2):
        p = f"quantizer.upsample.{i}.0.conv"
        # ConvTranspose1d weight layout [C_in, C_out, k]
        b.sd[p + ".weight"] = _randn(p + ".weight", seed, 512, 512, 2) / math.sqrt(512)
        b.bias(p + ".bias", 512)
        b.convnext(f"quantizer.upsample.{i}.1", 512)
    b.wn_conv("head.conv_pre.conv", (512, 512, 13), 512 * 13)
    b.bias("head.conv_pre.conv.bias", 512)
    chans = [512, 256, 128, 64, 32, 16]
    up_k = [16, 16, 4, 4, 4]
    up_s = [8, 8, 2, 2, 2]
    for i in range(5):
        # each output sample sees k/stride taps of every input channel
        b.wn_conv(f"head.ups.{i}.conv", (chans[i], chans[i + 1], up_k[i]), chans[i] * up_k[i] // up_s[i], gain=1.4)
        b.bias(f"head.ups.{i}.conv.bias", chans[i + 1])
        c = chans[i + 1]
        for j, k in enumerate((3, 7, 11)):
            for which in ("convs1", "convs2"):
                for d in range(3):
                    p = f"head.resblocks.{i}.blocks.{j}.{which}.{d}.conv"
                    b.wn_conv(p, (c, c, k), c * k, gain=1.0 if which == "convs1" else 0.6)
                    b.bias(p + ".bias", c)
    b.wn_conv("head.conv_post.conv", (1, 16, 13), 16 * 13, gain=0.35)
    b.bias("head.conv_post.conv.bias", 1, std=0.01)
    return b.sd


def make_vocoder_encoder_state_dict(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """The vocoder's ENCODE path (`FireflyArchitecture.encode`, modules/vqgan/modules/firefly.py:561-574: reference wave ->
    codec ids for the prompt, evaluations/infer_arvc.py:168-171): ConvNeXt backbone, quantizer.downsample and the FSQ
    project_in of each of the 8 groups.  Kept apart from make_vocoder_state_dict so that the decode-path fixtures and
    their weight digest do not change; merge the two dicts to get the full checkpoint."""
    b = _Builder(seed + 7)
    dims, depths = [128, 256, 384, 512], [3, 3, 9, 3]
    b.linear("backbone.downsample_layers.0.0.conv", dims[0], 160, bias=True, extra=(7,))
    b.norm("backbone.downsample_layers.0.1", dims[0], bias=True)
    for i in range(1, 4):
        b.norm(f"backbone.downsample_layers.{i}.0", dims[i - 1], bias=True)
        b.linear(f"backbone.downsample_layers.{i}.1", dims[i], dims[i - 1], bias=True, extra=(1,))
    for s in range(4):
        for j in range(depths[s]):
            b.convnext(f"backbone.stages.{s}.{j}", dims[s])
    b.norm("backbone.norm", 512, bias=True)
    for i in range(2):
        b.linear(f"quantizer.downsample.{i}.0.conv", 512, 512, bias=True, extra=(2,))
        b.convnext(f"quantizer.downsample.{i}.1", 512)
    for g in range(8):
        b.linear(f"quantizer.residual_fsq.rvqs.{g}.project_in", 4, 64, gain=3.0, bias=True)
    return b.sd


def make_campplus_state_dict(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Style encoder checkpoint (`pretrained_checkpoints/campplus_cn_common.bin`, configs/hydra_arcs/sv/campplus.yaml:
    CAMPPlus(feat_dim=80, embedding_size=192)); key names and shapes per modules/campplus/DTDNN.py:13-112 and
    modules/campplus/layers.py:10-266.  BatchNorm running statistics are non-trivial (eval-mode affine maps)."""
    b = _Builder(seed + 11)

    def bn(key, dim, affine=True):
        if affine:
            b.sd[key + ".weight"] = 1.0 + 0.1 * _randn(key + ".weight", b.seed, dim)
            b.sd[key + ".bias"] = 0.05 * _randn(key + ".bias", b.seed, dim)
        b.sd[key + ".running_mean"] = 0.1 * _randn(key + ".running_mean", b.seed, dim)
        b.sd[key + ".running_var"] = 1.0 + 0.2 * _randn(key + ".running_var", b.seed, dim).abs()
        b.sd[key + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.long)

    def conv2d(key, out_c, in_c, k):
        b.sd[key + ".weight"] = _randn(key + ".weight", b.seed, out_c, in_c, k, k) * (1.4 / math.sqrt(in_c * k * k))

    # FCM head (DTDNN.py:13-48): conv1/bn1, two stages of two BasicResBlocks (layers.py:223-266), conv2/bn2
    conv2d("head.conv1", 32, 1, 3)
    bn("head.bn1", 32)
    for layer in (1, 2):
        for blk in (0, 1):
            pre = f"head.layer{layer}.{blk}"
            conv2d(pre + ".conv1", 32, 32, 3)
            bn(pre + ".bn1", 32)
            conv2d(pre + ".conv2", 32, 32, 3)
            bn(pre + ".bn2", 32)
            if blk == 0:                                   # stride (2, 1): 1x1 conv + BN shortcut
                conv2d(pre + ".shortcut.0", 32, 32, 1)
                bn(pre + ".shortcut.1", 32)
    conv2d("head.conv2", 32, 32, 3)
    bn("head.bn2", 32)
    # x-vector trunk (DTDNN.py:63-95)
    b.linear("xvector.tdnn.linear", 128, 320, gain=1.4, extra=(5,))
    bn("xvector.tdnn.nonlinear.batchnorm", 128)
    channels = 128
    for i, (num_layers, _k, _d) in enumerate(zip((12, 24, 16), (3, 3, 3), (1, 2, 2))):
        for j in range(num_layers):
            pre = f"xvector.block{i + 1}.tdnnd{j + 1}"
            cin = channels + j * 32
            bn(pre + ".nonlinear1.batchnorm", cin)
            b.linear(pre + ".linear1", 128, cin, gain=1.4, extra=(1,))
            bn(pre + ".nonlinear2.batchnorm", 128)
            b.linear(pre + ".cam_layer.linear_local", 32, 128, gain=1.4, extra=(3,))
            b.linear(pre + ".cam_layer.linear1", 64, 128, gain=1.4, bias=True, extra=(1,))
            b.linear(pre + ".cam_layer.linear2", 32, 64, gain=2.0, bias=True, extra=(1,))
        channels += num_layers * 32
        bn(f"xvector.transit{i + 1}.nonlinear.batchnorm", channels)
        b.linear(f"xvector.transit{i + 1}.linear", channels // 2, channels, gain=1.4, extra=(1,))
        channels //= 2
    bn("xvector.out_nonlinear.batchnorm", channels)
    b.linear("dense.linear", 192, 2 * channels, extra=(1,))
    bn("dense.nonlinear.batchnorm", 192, affine=False)
    return b.sd


def make_timbre_encoder_state_dict(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Timbre encoder checkpoint (`pretrained_checkpoints/spark_speaker_encoder.pth`,
    configs/hydra_arcs/sv/sparktts_speaker_encoder.yaml), only the tensors `tokenize_wav` touches
    (modules/bicodec_speaker_encoder/speaker_encoder.py:136-144): ECAPA-TDNN trunk up to the 1536-channel latent
    (ecapa_tdnn.py:150-209; its pooling head is not on this path), PerceiverResampler (perceiver_encoder.py:300-351) and
    the FSQ projections (fsq/residual_fsq.py:70-76)."""
    b = _Builder(seed + 13)

    def bn(key, dim):
        b.sd[key + ".weight"] = 1.0 + 0.1 * _randn(key + ".weight", b.seed, dim)
        b.sd[key + ".bias"] = 0.05 * _randn(key + ".bias", b.seed, dim)
        b.sd[key + ".running_mean"] = 0.1 * _randn(key + ".running_mean", b.seed, dim)
        b.sd[key + ".running_var"] = 1.0 + 0.2 * _randn(key + ".running_var", b.seed, dim).abs()
        b.sd[key + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.long)

    e = "speaker_encoder"
    b.linear(e + ".layer1.conv", 512, 128, gain=0.3, bias=True, extra=(5,))     # mel magnitudes are O(1..30)
    bn(e + ".layer1.bn", 512)
    for layer in (2, 3, 4):
        pre = f"{e}.layer{layer}.se_res2block"
        b.linear(pre + ".0.conv", 512, 512, gain=1.4, bias=True, extra=(1,))
        bn(pre + ".0.bn", 512)
        for i in range(7):
            b.linear(f"{pre}.1.convs.{i}", 64, 64, gain=1.4, bias=True, extra=(3,))
            bn(f"{pre}.1.bns.{i}", 64)
        b.linear(pre + ".2.conv", 512, 512, gain=1.4, bias=True, extra=(1,))
        bn(pre + ".2.bn", 512)
        b.linear(pre + ".3.linear1", 128, 512, gain=1.4, bias=True)
        b.linear(pre + ".3.linear2", 512, 128, gain=2.0, bias=True)
    b.linear(e + ".conv", 1536, 1536, gain=1.4, bias=True, extra=(1,))
    ps = "perceiver_sampler"
    b.table(ps + ".latents", 32, 128, 0.5)
    b.linear(ps + ".proj_context", 128, 1536, bias=True)
    for layer in range(2):
        b.linear(f"{ps}.layers.{layer}.0.to_q", 512, 128, gain=1.5)
        b.linear(f"{ps}.layers.{layer}.0.to_kv", 1024, 128, gain=1.5)
        b.linear(f"{ps}.layers.{layer}.0.to_out", 128, 512, gain=0.7)
        b.linear(f"{ps}.layers.{layer}.1.0", 682, 128, bias=True)
        b.linear(f"{ps}.layers.{layer}.1.2", 128, 341, gain=0.7, bias=True)
    b.sd[ps + ".norm.gamma"] = 1.0 + 0.1 * _randn(ps + ".norm.gamma", b.seed, 128)
    b.linear("quantizer.project_in", 6, 128, gain=1.0, bias=True)
    b.linear("quantizer.project_out", 128, 6, gain=1.0, bias=True)
    return b.sd


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """w = g * v / ||v|| over dims (1,2) -- what `remove_parametrizations()` leaves behind
    (evaluations/infer_arvc.py:94, modules/vqgan/modules/firefly.py:105-111)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".parametrizations.weight.original1"):
            base = k[: -len(".parametrizations.weight.original1")]
            g = sd[base + ".parametrizations.weight.original0"]
            out[base + ".weight"] = g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)
        elif k.endswith(".parametrizations.weight.original0"):
            continue
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------- audio


def synth_audio_16k(seed: int, seconds: float) -> torch.Tensor:
    """Harmonic-stack test signal at 16 kHz: five harmonics of a
    random-walk f0 in 100-300 Hz, 4 Hz amplitude modulation, -40 dB noise floor."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    n = int(round(seconds * 16000))
    steps = torch.randn(n, generator=g) * 0.02
    f0 = 200.0 + torch.cumsum(steps, 0)
    f0 = 200.0 + 100.0 * torch.sin((f0 - 200.0) / 100.0)          # fold into 100..300 Hz
    phase = 2 * math.pi * torch.cumsum(f0 / 16000.0, 0)
    t = torch.arange(n) / 16000.0
    x = sum((1.0 / h) * torch.sin(h * phase) for h in range(1, 6))
    x = 0.3 * x * (0.5 + 0.5 * torch.sin(2 * math.pi * 4 * t)) + 0.01 * torch.randn(n, generator=g)
    return x.float()


def synth_audio_44k(seed: int, seconds: float) -> torch.Tensor:
    """16 kHz synthetic audio resampled to the model rate exactly as the host boundary does
    (`librosa.load(path, sr=44100)`, evaluations/infer_arvc.py:274,615)."""
    import torchaudio.functional as AF
    return AF.resample(synth_audio_16k(seed, seconds)[None], 16000, MODEL_SR)[0].contiguous()


def synth_speaker(seed: int):
    """Stand-ins for the two speaker-encoder outputs (setup path):
    style_vectors [1,192] (CAMPPlus) and timbre_latents [1,32,128] (SparkTTS)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(1, 192, generator=g), torch.randn(1, 32, 128, generator=g)


def noise_tape(seed: int, step: int, n_slots: int = 9, width: int = 8192) -> torch.Tensor:
    """Slot-indexed Exp(1) sampling noise for one `decode_one_token_ar` call
    (modules/dual_ar_stream.py:1092-1096 draws `q ~ Exp(1)` nine times per call: slot 0
    over 8192 entries, slots 1..8 over 1000).  Returned shape [9, 8192]; slot s uses
    the first V_s entries of its row."""
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + step * 7919 + 17) & 0x7FFFFFFF)
    u = torch.rand(n_slots, width, generator=g, dtype=torch.float64)
    return (-torch.log1p(-u)).clamp_min(1e-30).float()

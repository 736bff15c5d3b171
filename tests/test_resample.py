"""Host audio boundary: the GPU resampler against torchaudio.functional.resample, the function
the reference GUI and the synthetic-input harness use.  Tolerance: fp32 FIR sums of <= 475 taps on |x| <= 1 differ only
by summation order -> max abs error 5e-6."""
import math

import pytest
import torch

torchaudio = pytest.importorskip("torchaudio")

RATES = [(16000, 44100), (44100, 16000), (48000, 44100), (22050, 44100)]


@pytest.mark.parametrize("orig,new", RATES)
def test_filter_bank_is_torchaudios(orig, new):
    """CPU: the restated kernel construction gives bit-identical filter banks."""
    from torchaudio.functional import functional as Fn
    from streamvoiceanon_b200.audio import sinc_resample_kernel
    k, w = Fn._get_sinc_resample_kernel(orig, new, math.gcd(orig, new))
    mine, w2, o, n = sinc_resample_kernel(orig, new)
    assert w == w2 and (o, n) == (orig // math.gcd(orig, new), new // math.gcd(orig, new))
    assert torch.equal(k[:, 0], mine)


@pytest.mark.gpu
@pytest.mark.parametrize("orig,new", RATES)
def test_resampler_vs_torchaudio(orig, new):
    from streamvoiceanon_b200 import synth
    from streamvoiceanon_b200.audio import Resampler
    x = synth.synth_audio_16k(1600, 1.3)[: 20001]                      # odd length: partial last frame
    want = torchaudio.functional.resample(x, orig, new)
    rs = Resampler(orig, new)
    got_dev = rs(x.cuda())
    got_host = rs(x)
    assert got_dev.is_cuda and not got_host.is_cuda
    assert got_dev.shape == want.shape == got_host.shape
    assert float((got_dev.cpu() - want).abs().max()) < 5e-6
    assert torch.equal(got_dev.cpu(), got_host)


@pytest.mark.gpu
def test_resampler_batch_and_harness_equivalence():
    """[rows, n] input, and the bench harness' own 16 kHz -> 44.1 kHz step (synth.synth_audio_44k)."""
    from streamvoiceanon_b200 import synth
    from streamvoiceanon_b200.audio import Resampler
    x16 = torch.stack([synth.synth_audio_16k(1000 + i, 0.5) for i in range(2)])
    got = Resampler(16000, 44100)(x16.cuda()).cpu()
    want = torch.stack([synth.synth_audio_44k(1000 + i, 0.5) for i in range(2)])
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 5e-6

"""The caller-side `torch.compile(..., fullgraph=True)` of the reference (evaluations/infer_arvc.py:128-142,
real-time-gui.py:54-57) must keep working on the drop-in objects: the engine entries are custom ops with fake
implementations (streamvoiceanon_b200/ops.py), and the C ABI is CUDA-graph-capture safe (stream-ordered, no allocation
after warm-up, no host sync for device buffers)."""
import pytest
import torch

from streamvoiceanon_b200 import synth

pytestmark = pytest.mark.gpu


def _window():
    win = torch.zeros(1, 32 * 2048)
    win[:, -20 * 2048:] = synth.synth_audio_44k(4100, 1.0)[: 20 * 2048]
    return win.cuda()


def test_compiled_encode_and_head_fullgraph(models):
    """What infer_arvc.py:128-142 does, with the CPU-safe backend the reference itself selects (`aot_eager`)."""
    _, tok, voc = models
    wav, lens = _window(), torch.LongTensor([32 * 2048]).cuda()
    want_ids, want_len = tok.encode(wav, lens)
    enc = torch.compile(tok.encode, fullgraph=True, backend="aot_eager")
    ids, flen = enc(wav, lens)
    assert torch.equal(ids, want_ids) and torch.equal(flen, want_len)
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(0, 1000, (1, 8, 6), generator=g).cuda()
    z = voc.quantizer.decode(codes)
    want = voc.head(z)
    head = torch.compile(voc.head, fullgraph=True, backend="aot_eager")
    got = head(z)
    assert torch.equal(got, want)

    def code2wav(c):                               # infer_arvc.py:173-176
        return voc.head(voc.quantizer.decode(c))
    got2 = torch.compile(code2wav, fullgraph=True, backend="aot_eager")(codes)
    assert torch.equal(got2, want)


def test_entries_are_cuda_graph_capturable(models):
    """`mode="reduce-overhead"` captures the launches into a CUDA graph: capture and replay the encoder window and the
    vocoder head directly, results identical to the eager calls, also after the inputs change in place."""
    from streamvoiceanon_b200 import ops
    _, tok, voc = models
    wav = _window()
    g = torch.Generator().manual_seed(6)
    z = voc.quantizer.decode(torch.randint(0, 1000, (1, 8, 5), generator=g).cuda()).contiguous()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                     # warm-up on the side stream (workspace growth happens here)
        for _ in range(2):
            ops.enc_encode(wav)
            ops.voc_head(z)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ids = ops.enc_encode(wav)
        wave = ops.voc_head(z)
    for trial in range(2):
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(ids, tok.encode(wav, torch.LongTensor([wav.shape[1]]).cuda())[0])
        assert torch.equal(wave, voc.head(z))
        wav.copy_(torch.roll(wav, 2048 * 3, dims=1))            # new contents, same static buffers
        z.mul_(0.9)


def test_compiled_reduce_overhead(models):
    """The reference's exact call: inductor + CUDA graphs around the custom op."""
    _, tok, _ = models
    wav, lens = _window(), torch.LongTensor([32 * 2048]).cuda()
    want, _ = tok.encode(wav, lens)
    enc = torch.compile(tok.encode, fullgraph=True, mode="reduce-overhead")
    for _ in range(3):
        ids, _ = enc(wav, lens)
    assert torch.equal(ids.clone(), want)

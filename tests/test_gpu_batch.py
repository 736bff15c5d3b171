"""GPU tests of the many-stream (lock-step) path.  The reference is batch-1, so the contract is per-stream parity:
stream i of a batch == the same stream alone (which tests/test_gpu_parity.py pins to the reference fixtures and the
oracle), plus one direct check of a batched stream against the reference fixture.  Integer outputs bit-exact;
waveforms within the fp32 MSE tolerance written beside each check (GEMM tile/split choices differ with M)."""
import ctypes as C

import numpy as np
import pytest
import torch

from streamvoiceanon_b200 import synth

pytestmark = pytest.mark.gpu

WAVE_MSE_TOL = 1e-8


def test_encode_batch_equals_single(models):
    _, tok, _ = models
    n = 21 * 2048
    wavs = torch.stack([synth.synth_audio_44k(3000 + i, 1.2)[:n] for i in range(3)]).cuda()
    lens = torch.LongTensor([n] * 3).cuda()
    both, flen = tok.encode(wavs, lens)
    assert tuple(both.shape) == (1, 3, 21) and flen.tolist() == [21, 21, 21]
    for i in range(3):
        one, _ = tok.encode(wavs[i:i + 1].contiguous(), lens[:1])
        assert torch.equal(one[0, 0], both[0, i]), i


def test_encode_batch_many_rows(models):
    """>= 8 rows switch the window attention to one CTA per (stream, head); 33 frames: ragged last query block."""
    _, tok, _ = models
    n = 33 * 2048
    wavs = torch.stack([synth.synth_audio_44k(3200 + i, 1.7)[:n] for i in range(9)]).cuda()
    lens = torch.LongTensor([n] * 9).cuda()
    both, _ = tok.encode(wavs, lens)
    for i in (0, 4, 8):
        one, _ = tok.encode(wavs[i:i + 1].contiguous(), lens[:1])
        assert torch.equal(one[0, 0], both[0, i]), i


def test_encode_batch_streaming_window_vs_reference(models, gold):
    """Row 1 of a 2-row batched call is the reference's 128-frame streaming-window fixture."""
    _, tok, _ = models
    g = gold("encoder_window128")
    live = int(g["live_frames"])
    win = torch.zeros(2, 128 * 2048)
    win[0] = synth.synth_audio_44k(3100, 6.5)[: 128 * 2048]
    win[1, -live * 2048:] = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: live * 2048]
    ids, _ = tok.encode(win.cuda(), torch.LongTensor([win.shape[1]] * 2).cuda())
    assert np.array_equal(ids[:, 1:2].cpu().numpy(), g["ids"])


def _prompts(n, seed=900):
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(n):
        T = 20 + 9 * b
        out.append((torch.randint(0, 8192, (1, T), generator=g), torch.randint(0, 1000, (1, 8, T), generator=g).int(),
                    torch.randint(0, 8192, (1, 8), generator=g), synth.synth_speaker(6000 + b)))
    return out


@pytest.mark.parametrize("n", [1, 3, 5])
def test_decode_many_equals_single_stream(models, tape, n):
    """svanon_ar_decode_many (GEMM path, any stream count) == the persistent single-stream kernel, codes bit-exact."""
    from streamvoiceanon_b200 import ARVCWrapper, _lib
    from streamvoiceanon_b200.engine import ptr
    ar0, _, _ = models
    prompts = _prompts(n)
    singles = []
    for b, (rc, ra, src, (style, timbre)) in enumerate(prompts):
        ar0.set_delay(delay=2)
        ar0.set_noise_fn(tape(7200 + b), 0)
        ar0.prefill_prompt(rc.cuda(), ra.cuda(), style.cuda(), timbre.cuda())
        ar0.prefill_src_condition4delay(src[:, :2].cuda())
        singles.append([ar0.decode_one(src[:, t:t + 1].cuda())[0].cpu() for t in range(2, 8)])
    wrappers = []
    for b, (rc, ra, src, (style, timbre)) in enumerate(prompts):
        w = ARVCWrapper()
        w.setup_caches(max_batch_size=1, max_seq_len=2048)
        w.set_delay(delay=2)
        w.prefill_prompt(rc.cuda(), ra.cuda(), style.cuda(), timbre.cuda())
        w.prefill_src_condition4delay(src[:, :2].cuda())
        wrappers.append(w)
    lib = _lib.load()
    handles = (C.c_void_p * n)(*[w._stream for w in wrappers])
    for i, t in enumerate(range(2, 8)):
        ids = torch.tensor([int(p[2][0, t]) for p in prompts], dtype=torch.int64).cuda()
        noise = torch.stack([torch.stack([tape(7200 + b)(2 + i, s, 1000)[:1000] for s in range(1, 9)]) for b in range(n)])
        noise = noise.float().contiguous().cuda()
        out = torch.empty(n, 8, dtype=torch.int32, device="cuda")
        _lib.check(lib.svanon_ar_decode_many(handles, n, ptr(ids), ptr(noise), ptr(out), None))
        for b in range(n):
            assert torch.equal(out[b].cpu(), singles[b][i][:, 0]), (b, i)
            assert int(lib.svanon_ar_position(wrappers[b]._stream)) == 33 + 2 * prompts[b][0].shape[1] + 3 + 2 * (i + 1)


def _stream_inputs(tok, b, n_ref, n_chunks, chunk):
    style, timbre = synth.synth_speaker(5100 + b)
    ref_wave = synth.synth_audio_44k(5100 + b, 3.5)[: n_ref * 2048][None]
    gen = torch.Generator().manual_seed(300 + b)
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    src = synth.synth_audio_44k(1100 + b, 3.0)[: n_chunks * chunk * 2048].view(n_chunks, chunk * 2048)
    return ref_content[0], ref_audio, style, timbre, src


def _session(inp, tape_fn, delay):
    from streamvoiceanon_b200 import StreamSession
    ref_content, ref_audio, style, timbre, _ = inp
    sess = StreamSession()
    sess.set_noise_fn(tape_fn, 0)
    sess.set_prompt(ref_content.cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda(), max_prompt_frames=256, delay=delay)
    return sess


@pytest.mark.parametrize("n,chunk,ar_path,enc_mode", [(3, 1, 0, 1), (2, 2, 1, 1), (2, 1, 0, 1)])
def test_batch_loop_equals_single_sessions(models, tape, n, chunk, ar_path, enc_mode):
    """N streams with different prompts (lengths 26, 31, ...), sources and noise tapes, small windows so that the
    streams re-prompt at DIFFERENT chunks (stream 0 after 9 frames, stream 1 after 4, stream 2 at once): ids bit-exact and waveform equal to fp32 rounding vs each stream alone."""
    from streamvoiceanon_b200 import BatchSession
    _, tok, _ = models
    n_chunks, delay = 16, 2
    cfg = dict(encode_window_frames=24, decode_window_frames=24, max_seq_frames=52, buffer_frames=6,
               decode_chunk_frames=chunk)
    inputs = [_stream_inputs(tok, b, 26 + 5 * b, n_chunks, chunk) for b in range(n)]
    singles = []
    for b, inp in enumerate(inputs):
        sess = _session(inp, tape(7300 + b), delay)
        sess.setup(**cfg)
        waves = torch.cat([sess.process_chunk(inp[4][i].cuda()).cpu() for i in range(n_chunks)])
        singles.append((*sess.history(), waves))
        sess.close()
    sessions = [_session(inp, tape(7300 + b), delay) for b, inp in enumerate(inputs)]
    batch = BatchSession(sessions)
    batch.setup(**cfg)
    batch.set_ar_path(ar_path)
    outs = []
    for i in range(n_chunks):
        w = torch.stack([inp[4][i] for inp in inputs])
        outs.append(batch.process_chunk(w.cuda() if i % 2 == 0 else w).cpu())       # device and host buffers
    waves = torch.cat(outs, dim=1)
    for b, sess in enumerate(sessions):
        src_hist, pred_hist = sess.history()
        assert torch.equal(src_hist, singles[b][0]), b
        assert torch.equal(pred_hist, singles[b][1]), b
        mse = float(((waves[b] - singles[b][2]) ** 2).mean())
        assert mse < 1e-10, (b, mse)
    batch.close()
    for s in sessions:
        s.close()


def test_stream_pool_equals_single_sessions(models, tape):
    """StreamPool (streams joining and leaving at chunk boundaries, two cohorts alive at once): every stream produces
    what it produces alone -- ids bit-exact, waveform to fp32 rounding."""
    from streamvoiceanon_b200.server import StreamPool
    _, tok, _ = models
    n_chunks, delay = 12, 2
    cfg = dict(encode_window_frames=24, decode_window_frames=24, max_seq_frames=52, buffer_frames=6, decode_chunk_frames=1)
    inputs = [_stream_inputs(tok, b, 26 + 5 * b, n_chunks, 1) for b in range(3)]
    singles = []
    for b, inp in enumerate(inputs):
        sess = _session(inp, tape(7400 + b), delay)
        sess.setup(**cfg)
        waves = torch.cat([sess.process_chunk(inp[4][i].cuda()).cpu() for i in range(n_chunks)])
        singles.append((*sess.history(), waves))
        sess.close()
    sessions = [_session(inp, tape(7400 + b), delay) for b, inp in enumerate(inputs)]
    pool = StreamPool(merge_cohorts=False, **cfg)
    join = {0: 0, 1: 0, 2: 3}                       # stream 2 arrives three chunks later -> its own cohort
    got = {b: [] for b in range(3)}
    for step in range(n_chunks + 3):
        for b, at in join.items():
            if at == step:
                pool.add(b, sessions[b])
        chunks = {b: inputs[b][4][step - join[b]].cuda() for b in range(3) if b in pool and step - join[b] < n_chunks}
        for b, w in pool.step(chunks).items():
            if b in chunks:
                got[b].append(w.cpu())
        if step == n_chunks - 1:
            assert pool.n_cohorts == 2 and pool.cohort_sizes() == [2, 1]
            pool.remove(0)                          # streams 0 and 1 are done; 1 leaves last and closes the cohort
            pool.remove(1)
            assert pool.n_cohorts == 1
    for b, sess in enumerate(sessions):
        src_hist, pred_hist = sess.history()
        assert torch.equal(src_hist[: singles[b][0].numel()], singles[b][0]), b
        assert torch.equal(pred_hist[:, : singles[b][1].shape[1]], singles[b][1]), b
        mse = float(((torch.cat(got[b]) - singles[b][2]) ** 2).mean())
        assert mse < 1e-10, (b, mse)
    pool.close()
    for s in sessions:
        s.close()


@pytest.mark.parametrize("enc_mode", [1, 3])
def test_stream_pool_merges_cohorts(models, tape, enc_mode):
    """Cohort merging (svanon_batch_merge): three streams arrive at chunks 0, 0 and 3; once the late cohort has left its
    warm-up it is folded into the first one -- wave rings, encoder window state (enc_mode 1) or stateful-encoder state with
    its K/V rings (enc_mode 3), vocoder histories and history columns move over -- and every stream still produces what it
    produces alone: ids bit-exact, waveform to fp32 rounding.  Small windows, so that re-prompts fire after the merge."""
    from streamvoiceanon_b200 import BatchSession
    from streamvoiceanon_b200.server import StreamPool
    _, tok, _ = models
    n_chunks, delay = 16, 2
    cfg = dict(encode_window_frames=24, decode_window_frames=24, max_seq_frames=52, buffer_frames=6, decode_chunk_frames=1)
    inputs = [_stream_inputs(tok, b, 26 + 5 * b, n_chunks, 1) for b in range(3)]
    singles = []
    for b, inp in enumerate(inputs):
        sess = _session(inp, tape(7450 + b), delay)
        sess.set_encoder_mode(enc_mode)
        sess.setup(**cfg)
        waves = torch.cat([sess.process_chunk(inp[4][i].cuda()).cpu() for i in range(n_chunks)])
        singles.append((*sess.history(), waves))
        sess.close()
    sessions = [_session(inp, tape(7450 + b), delay) for b, inp in enumerate(inputs)]

    class ModeBatch(BatchSession):
        def setup(self, **kw):
            self.set_encoder_mode(enc_mode)
            super().setup(**kw)
    pool = StreamPool(batch_factory=ModeBatch, **cfg)
    join = {0: 0, 1: 0, 2: 3}
    got = {b: [] for b in range(3)}
    sizes = []
    for step in range(n_chunks + 3):
        for b, at in join.items():
            if at == step:
                pool.add(b, sessions[b])
        chunks = {b: inputs[b][4][step - join[b]].cuda() for b in range(3) if b in pool and step - join[b] < n_chunks}
        for b, w in pool.step(chunks).items():
            if b in chunks:
                got[b].append(w.cpu())
        sizes.append(pool.cohort_sizes())
    assert [2, 1] in sizes and sizes[-1] == [3] and pool.merges == 1          # two cohorts for a while, then one
    for b, sess in enumerate(sessions):
        src_hist, pred_hist = sess.history()
        assert torch.equal(src_hist[: singles[b][0].numel()], singles[b][0]), b
        assert torch.equal(pred_hist[:, : singles[b][1].shape[1]], singles[b][1]), b
        mse = float(((torch.cat(got[b]) - singles[b][2]) ** 2).mean())
        assert mse < 1e-10, (b, mse)
    pool.close()
    for s in sessions:
        s.close()


def test_stream_pool_merge_then_long_run(models, tape):
    """Two cohorts of 8 streams (CLI-default windows), the second arriving 5 chunks late and folded into the first after its
    warm-up: the merged batch starts with EMPTY rings of steady-state layer inputs (ConvStackRings), runs the whole-span
    window-start pass until the rings cover the 128-frame window again (chunk ~136), and the ring-based pass after that.
    200 chunks; one stream of each cohort against the same stream run alone: ids bit-exact."""
    from streamvoiceanon_b200 import BatchSession
    from streamvoiceanon_b200.server import StreamPool
    _, tok, _ = models
    n_chunks, late, delay = 200, 5, 2
    cfg = dict(encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32, decode_chunk_frames=1)

    def inputs_of(b):
        ref_content, ref_audio, style, timbre, _ = _stream_inputs(tok, 80 + b, 60 + 3 * b, 4, 1)
        src = synth.synth_audio_44k(1700 + b, 10.0)[: n_chunks * 2048].view(n_chunks, 2048)
        return ref_content, ref_audio, style, timbre, src
    inputs = [inputs_of(b) for b in range(16)]
    check = (3, 12)
    singles = {}
    for b in check:
        sess = _session(inputs[b], tape(6600 + b), delay)
        sess.setup(**cfg)
        for i in range(n_chunks):
            sess.process_chunk(inputs[b][4][i].cuda())
        singles[b] = sess.history()
        sess.close()
    sessions = [_session(inp, tape(6600 + b), delay) for b, inp in enumerate(inputs)]
    pool = StreamPool(batch_factory=BatchSession, **cfg)
    join = {b: (0 if b < 8 else late) for b in range(16)}
    sizes = []
    for step in range(n_chunks + late):
        for b, at in join.items():
            if at == step:
                pool.add(b, sessions[b])
        chunks = {b: inputs[b][4][step - join[b]].cuda() for b in range(16) if b in pool and step - join[b] < n_chunks}
        pool.step(chunks)
        sizes.append(pool.cohort_sizes())
    assert [8, 8] in sizes and [16] in sizes and pool.merges == 1
    for b in check:
        src_hist, pred_hist = sessions[b].history()
        assert torch.equal(src_hist[: singles[b][0].numel()], singles[b][0]), b
        assert torch.equal(pred_hist[:, : singles[b][1].shape[1]], singles[b][1]), b
    pool.close()
    for s in sessions:
        s.close()


@pytest.mark.parametrize("enc_mode", [1, 3])
def test_stream_pool_drops_streams_that_left(models, tape, enc_mode):
    """Compaction (svanon_batch_select): streams 0-2 arrive at chunk 0, stream 3 three chunks later (merged into the first
    cohort once warm, so its history offsets are non-zero); stream 1 leaves after 8 chunks and stream 0 after 10.  Each time the
    silent members reach a quarter of the cohort it continues with the live members only -- members [0, 2, 3], then [2, 3], at
    the end [3] alone: wave
    rings, encoder state (enc_mode 3: stateful K/V rings and positions), vocoder histories and history columns gathered per
    member.  Every stream equals the stream alone for as long as it ran: ids bit-exact, waveform to fp32 rounding."""
    from streamvoiceanon_b200 import BatchSession
    from streamvoiceanon_b200.server import StreamPool
    _, tok, _ = models
    n_chunks, delay = 16, 2
    cfg = dict(encode_window_frames=24, decode_window_frames=24, max_seq_frames=52, buffer_frames=6, decode_chunk_frames=1)
    inputs = [_stream_inputs(tok, b, 26 + 5 * b, n_chunks, 1) for b in range(4)]
    join = {0: 0, 1: 0, 2: 0, 3: 3}
    runs = {0: 10, 1: 8, 2: n_chunks, 3: n_chunks}
    singles = []
    for b, inp in enumerate(inputs):
        sess = _session(inp, tape(7480 + b), delay)
        sess.set_encoder_mode(enc_mode)
        sess.setup(**cfg)
        waves = torch.cat([sess.process_chunk(inp[4][i].cuda()).cpu() for i in range(runs[b])])
        singles.append((*sess.history(), waves))
        sess.close()
    sessions = [_session(inp, tape(7480 + b), delay) for b, inp in enumerate(inputs)]

    class ModeBatch(BatchSession):
        def setup(self, **kw):
            self.set_encoder_mode(enc_mode)
            super().setup(**kw)
    pool = StreamPool(batch_factory=ModeBatch, compact_fraction=0.25, **cfg)
    got = {b: [] for b in range(4)}
    sizes = []
    for step in range(n_chunks + 3):
        for b, at in join.items():
            if at == step:
                pool.add(b, sessions[b])
        for b in range(4):
            if b in pool and step - join[b] == runs[b]:
                pool.remove(b)
        chunks = {b: inputs[b][4][step - join[b]].cuda() for b in range(4) if b in pool}
        for b, w in pool.step(chunks).items():
            got[b].append(w.cpu())
        sizes.append(pool.cohort_sizes())
    # merged at chunk 7; compacted when stream 1, stream 0 and (after its 16 chunks) stream 2 had left: stream 3 ends alone
    assert pool.merges == 1 and pool.compactions == 3 and all(z in sizes for z in ([3, 1], [4], [3], [2])) and sizes[-1] == [1], sizes
    assert not any(pool.in_use(sessions[b]) for b in (0, 1, 2)) and pool.in_use(sessions[3])
    for b, sess in enumerate(sessions):
        src_hist, pred_hist = sess.history()
        assert torch.equal(src_hist[: singles[b][0].numel()], singles[b][0]), b
        assert torch.equal(pred_hist[:, : singles[b][1].shape[1]], singles[b][1]), b
        assert len(got[b]) == runs[b]
        mse = float(((torch.cat(got[b]) - singles[b][2]) ** 2).mean())
        assert mse < 1e-10, (b, mse)
    pool.close()
    for s in sessions:
        s.close()


def test_batch_select_argument_errors(models, tape):
    """svanon_batch_select refuses a batch that is still in its warm-up chunks and member lists that are not strictly
    increasing indices; the batch stays usable after a refused call."""
    from streamvoiceanon_b200 import BatchSession
    _, tok, _ = models
    cfg = dict(encode_window_frames=24, decode_window_frames=24, max_seq_frames=52, buffer_frames=6, decode_chunk_frames=1)
    inputs = [_stream_inputs(tok, b, 26 + 5 * b, 6, 1) for b in range(2)]
    sessions = [_session(inp, tape(7490 + b), 2) for b, inp in enumerate(inputs)]
    batch = BatchSession(sessions)
    batch.setup(**cfg)
    waves = lambda i: torch.stack([inp[4][i] for inp in inputs]).cuda()
    batch.process_chunk(waves(0))
    with pytest.raises(RuntimeError, match="warm-up"):
        BatchSession.selected(batch, [0])
    for i in range(1, 5):
        batch.process_chunk(waves(i))
    for bad in ([1, 0], [0, 0], [2], [-1], []):
        with pytest.raises(RuntimeError):
            BatchSession.selected(batch, bad)
    one = BatchSession.selected(batch, [1])
    assert len(one.sessions) == 1 and one.sessions[0] is sessions[1] and batch._h is None
    assert one.process_chunk(inputs[1][4][5][None].cuda()).shape == (1, 2048)
    one.close()
    for s in sessions:
        s.close()


@pytest.mark.parametrize("n,leave,n_chunks", [(10, 3, 40), (12, 4, 160)])
def test_stream_pool_compaction_default_windows(models, tape, n, leave, n_chunks):
    """Compaction with the CLI-default windows, where the cohort carries encoder window state: n streams from chunk 0, `leave`
    of them (every third member) gone after 12 chunks.  10 -> 7 members: the per-layer conv history mode (>= 8 streams) hands
    over to the tail-span mode on the gathered transformer inputs; 12 -> 8: the conv history itself is gathered and the rings
    of steady-state layer inputs start again empty and serve the window-start pass once they cover the window (chunk ~140 of
    160).  Two remaining streams against the same streams alone: ids bit-exact."""
    from streamvoiceanon_b200 import BatchSession
    from streamvoiceanon_b200.server import StreamPool
    _, tok, _ = models
    gone_at, delay = 12, 2
    cfg = dict(encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32, decode_chunk_frames=1)

    def inputs_of(b):
        ref_content, ref_audio, style, timbre, _ = _stream_inputs(tok, 60 + b, 64 + b, 4, 1)
        src = synth.synth_audio_44k(1900 + b, 8.0)[: n_chunks * 2048].view(n_chunks, 2048)
        return ref_content, ref_audio, style, timbre, src
    inputs = [inputs_of(b) for b in range(n)]
    leaving = [3 * i for i in range(leave)]                      # members 0, 3, 6(, 9): the rest is gathered in runs of two
    check = (1, n - 1)
    singles = {}
    for b in check:
        sess = _session(inputs[b], tape(6700 + b), delay)
        sess.setup(**cfg)
        for i in range(n_chunks):
            sess.process_chunk(inputs[b][4][i].cuda())
        singles[b] = sess.history()
        sess.close()
    sessions = [_session(inp, tape(6700 + b), delay) for b, inp in enumerate(inputs)]
    pool = StreamPool(batch_factory=BatchSession, compact_fraction=0.25, **cfg)
    for b in range(n):
        pool.add(b, sessions[b])
    for step in range(n_chunks):
        if step == gone_at:
            for b in leaving:
                pool.remove(b)
        pool.step({b: inputs[b][4][step].cuda() for b in range(n) if b in pool})
    assert pool.compactions == 1 and pool.cohort_sizes() == [n - leave]
    for b in check:
        src_hist, pred_hist = sessions[b].history()
        assert torch.equal(src_hist[: singles[b][0].numel()], singles[b][0]), b
        assert torch.equal(pred_hist[:, : singles[b][1].shape[1]], singles[b][1]), b
    pool.close()
    for s in sessions:
        s.close()


@pytest.mark.parametrize("enc_mode", [1, 2])
def test_batch_loop_vs_reference_fixture(models, gold, tape, enc_mode):
    """Stream 0 of a 2-stream batch (many-stream decode kernels forced) reproduces the UNMODIFIED reference's
    process_one_chunk run with CLI-default windows (tests/golden/stream_default.npz); enc_mode 1 = ring-buffer
    encoder state with a tail span, 2 = newest frames from per-layer conv history."""
    from streamvoiceanon_b200 import BatchSession, StreamSession
    _, tok, _ = models
    g = gold("stream_default")
    n_ref, n_chunks = int(g["n_ref"]), int(g["n_chunks"])
    style, timbre = synth.synth_speaker(int(g["ref_seed"]))
    ref_wave = synth.synth_audio_44k(int(g["ref_seed"]), 3.5)[: n_ref * 2048][None]
    gen = torch.Generator().manual_seed(int(g["codes_seed"]))
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    s0 = StreamSession()
    s0.set_noise_fn(tape(int(g["tape_seed"])), 0)
    s0.set_prompt(ref_content[0].cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda(), 256, int(g["delay"]))
    other = _stream_inputs(tok, 7, 40, n_chunks, 1)
    s1 = _session(other, tape(7400), int(g["delay"]))
    batch = BatchSession([s0, s1])
    batch.setup(int(g["encode_window_frames"]), int(g["decode_window_frames"]), int(g["max_seq_frames"]),
                int(g["buffer_frames"]), 1)
    batch.set_ar_path(1)
    batch.set_encoder_mode(enc_mode)
    src = synth.synth_audio_44k(int(g["src_seed"]), 1.5)[: n_chunks * 2048].view(n_chunks, 2048)
    waves = torch.cat([batch.process_chunk(torch.stack([src[i], other[4][i]]).cuda())[0].cpu() for i in range(n_chunks)])
    src_hist, pred_hist = s0.history()
    assert np.array_equal(src_hist.numpy()[None], g["src_content"])
    assert np.array_equal(pred_hist.numpy()[None], g["pred_codes"])
    mse = float(((waves.numpy() - g["wave"]) ** 2).mean())
    assert mse < WAVE_MSE_TOL, mse
    batch.close(); s0.close(); s1.close()


def test_batch_of_nine_default_windows(models, tape):
    """9 streams, CLI-default windows: the batch switches on its own to one attention CTA per (stream, head), the
    many-stream decode kernels and the per-layer conv history for the newest encoder frames.  Streams 0 and 8 against
    the same streams run alone: ids bit-exact, waveform to fp32 rounding."""
    from streamvoiceanon_b200 import BatchSession
    _, tok, _ = models
    n, n_chunks = 9, 10
    cfg = dict(encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32, decode_chunk_frames=1)
    inputs = [_stream_inputs(tok, 20 + b, 70 + 3 * b, n_chunks, 1) for b in range(n)]
    singles = {}
    for b in (0, 8):
        sess = _session(inputs[b], tape(7700 + b), 2)
        sess.setup(**cfg)
        waves = torch.cat([sess.process_chunk(inputs[b][4][i].cuda()).cpu() for i in range(n_chunks)])
        singles[b] = (*sess.history(), waves)
        sess.close()
    sessions = [_session(inp, tape(7700 + b), 2) for b, inp in enumerate(inputs)]
    batch = BatchSession(sessions)
    batch.setup(**cfg)
    waves = torch.cat([batch.process_chunk(torch.stack([inp[4][i] for inp in inputs]).cuda()).cpu() for i in range(n_chunks)], dim=1)
    for b in (0, 8):
        src_hist, pred_hist = sessions[b].history()
        assert torch.equal(src_hist, singles[b][0]), b
        assert torch.equal(pred_hist, singles[b][1]), b
        assert float(((waves[b] - singles[b][2]) ** 2).mean()) < 1e-10, b
    batch.close()
    for s in sessions:
        s.close()


@pytest.mark.parametrize("chunk,n_chunks", [(1, 300), (2, 150)])
def test_batch_long_run_window_start_rings(models, tape, chunk, n_chunks):
    """300 frames (as one- or two-frame chunks) of 8 lock-step streams with the CLI-default windows: long enough for the rings of steady-state layer inputs
    (ConvStackRings, 256 frames deep for a 128-frame window) to wrap, so the window-start pass that recomputes only the rows the
    zero padding reaches (Engine::enc_conv_stack_head; merged with the pass of the newest frames: enc_conv_stack_merged) reads
    rows written hundreds of chunks earlier.  Streams 0 and 7 against
    the same streams run alone (whole-span recompute, reference semantics): content ids and codec ids bit-exact."""
    from streamvoiceanon_b200 import BatchSession
    _, tok, _ = models
    n = 8
    cfg = dict(encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32, decode_chunk_frames=chunk)

    def inputs_of(b):
        ref_content, ref_audio, style, timbre, _ = _stream_inputs(tok, 60 + b, 64 + 5 * b, 4, 1)
        src = synth.synth_audio_44k(1500 + b, 14.5)[: n_chunks * chunk * 2048].view(n_chunks, chunk * 2048)
        return ref_content, ref_audio, style, timbre, src
    inputs = [inputs_of(b) for b in range(n)]
    singles = {}
    for b in (0, 7):
        sess = _session(inputs[b], tape(9900 + b), 2)
        sess.setup(**cfg)
        for i in range(n_chunks):
            sess.process_chunk(inputs[b][4][i].cuda())
        singles[b] = sess.history()
        sess.close()
    sessions = [_session(inp, tape(9900 + b), 2) for b, inp in enumerate(inputs)]
    batch = BatchSession(sessions)
    batch.setup(**cfg)
    for i in range(n_chunks):
        batch.process_chunk(torch.stack([inp[4][i] for inp in inputs]).cuda())
    for b in (0, 7):
        src_hist, pred_hist = sessions[b].history()
        assert src_hist.shape == singles[b][0].shape
        bad = (src_hist != singles[b][0]).nonzero()
        assert bad.numel() == 0, (b, "content ids differ first at", bad[:4].tolist())
        assert torch.equal(pred_hist, singles[b][1]), b
    batch.close()
    for s in sessions:
        s.close()


def test_batch_of_128_default_windows(models, tape):
    """BASELINE config 4's per-GPU share: 128 streams in lock-step with the CLI-default windows, so that the loop runs on
    the GEMM tiles the headline stream count uses (M = 128 x 512 encoder rows: the 128 x 256 / 128 x 128 tensor-core tiles,
    M = 256 decode rows) -- streams 0, 63 and 127 against the same streams run alone (which test_gpu_parity.py pins to
    the reference): content ids and codec ids bit-exact, waveform to fp32 rounding.  Every stream has its own prompt
    codes, speaker vectors, source and noise tape; four prompt lengths."""
    from streamvoiceanon_b200 import BatchSession
    _, tok, _ = models
    n, n_chunks, check = 128, 7, (0, 63, 127)
    cfg = dict(encode_window_frames=128, decode_window_frames=64, max_seq_frames=768, buffer_frames=32, decode_chunk_frames=1)
    base = [_stream_inputs(tok, 40 + k, 66 + 7 * k, n_chunks, 1) for k in range(4)]      # four distinct prompts / sources

    def inputs_of(b):
        ref_content, _, _, _, src = base[b % 4]
        gen = torch.Generator().manual_seed(8800 + b)
        ref_audio = torch.randint(0, 1000, (1, 8, ref_content.numel()), generator=gen).int()
        style, timbre = synth.synth_speaker(5300 + b)
        return ref_content, ref_audio, style, timbre, src.roll(b % 5, dims=0) * (0.5 + 0.5 * ((b * 37) % 11) / 10.0)
    singles = {}
    for b in check:
        inp = inputs_of(b)
        sess = _session(inp, tape(9100 + b), 2)
        sess.setup(**cfg)
        waves = torch.cat([sess.process_chunk(inp[4][i].cuda()).cpu() for i in range(n_chunks)])
        singles[b] = (*sess.history(), waves)
        sess.close()
    inputs = [inputs_of(b) for b in range(n)]
    sessions = [_session(inp, tape(9100 + b), 2) for b, inp in enumerate(inputs)]
    batch = BatchSession(sessions)
    batch.setup(**cfg)
    waves = torch.cat([batch.process_chunk(torch.stack([inp[4][i] for inp in inputs]).cuda()).cpu() for i in range(n_chunks)], dim=1)
    for b in check:
        src_hist, pred_hist = sessions[b].history()
        assert torch.equal(src_hist, singles[b][0]), b
        assert torch.equal(pred_hist, singles[b][1]), b
        assert float(((waves[b] - singles[b][2]) ** 2).mean()) < 1e-10, b
    batch.close()
    for s in sessions:
        s.close()


def test_batch_errors(models, tape):
    from streamvoiceanon_b200 import BatchSession
    _, tok, _ = models
    inp = _stream_inputs(tok, 0, 26, 2, 1)
    a, b = _session(inp, None, 2), _session(inp, None, 4)
    with pytest.raises(RuntimeError, match="same delay"):
        BatchSession([a, b]).setup(24, 24, 36, 6, 1)
    with pytest.raises(RuntimeError, match="incremental vocoder"):
        BatchSession([a]).setup(24, 8, 36, 6, 1)
    with pytest.raises(RuntimeError, match="only once"):
        BatchSession([a, a])
    a.close(); b.close()

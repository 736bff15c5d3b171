"""Host logic of StreamPool (streams joining / leaving at chunk boundaries, cohorts) with a stand-in batch object -- no GPU.
The GPU contract (every pooled stream equals the stream alone) is tests/test_gpu_batch.py::test_stream_pool_equals_single_sessions."""
import pytest
import torch

from streamvoiceanon_b200.server import StreamPool


class FakeSession:
    def __init__(self, gain, delay=2):
        self.gain, self.delay = gain, delay


class FakeBatch:
    """process_chunk: row i -> gain_i * wave + (number of chunks this batch has seen)."""
    log = []

    def __init__(self, sessions):
        self.sessions, self.n, self.closed, self.cfg = list(sessions), 0, False, None
        FakeBatch.log.append(self)

    def setup(self, **cfg):
        self.cfg = cfg

    def process_chunk(self, waves):
        assert not self.closed and waves.shape == (len(self.sessions), self.cfg["decode_chunk_frames"] * 2048)
        self.n += 1
        return torch.stack([s.gain * w + self.n for s, w in zip(self.sessions, waves)])

    def close(self):
        self.closed = True


@pytest.fixture
def pool():
    FakeBatch.log = []
    return StreamPool(encode_window_frames=24, decode_window_frames=24, max_seq_frames=52, buffer_frames=6,
                      decode_chunk_frames=1, batch_factory=FakeBatch)


def _chunk(v):
    return torch.full((2048,), float(v))


def test_cohorts_form_per_join_step_and_delay(pool):
    pool.add("a", FakeSession(1.0))
    pool.add("b", FakeSession(2.0))
    pool.add("c", FakeSession(3.0, delay=0))
    out = pool.step({"a": _chunk(1), "b": _chunk(1), "c": _chunk(1)})
    assert pool.n_cohorts == 2 and sorted(pool.cohort_sizes()) == [1, 2]          # delay 0 apart from delay 2
    assert float(out["a"][0]) == 2.0 and float(out["b"][0]) == 3.0 and float(out["c"][0]) == 4.0
    assert FakeBatch.log[0].cfg["encode_window_frames"] == 24
    pool.add("d", FakeSession(10.0))                                               # joins two chunks later: own cohort
    pool.step({"a": _chunk(0), "b": _chunk(0), "c": _chunk(0)})                    # d starts with the next step it sees
    assert pool.n_cohorts == 3
    out = pool.step({"a": _chunk(1), "b": _chunk(1), "c": _chunk(1), "d": _chunk(1)})
    assert float(out["a"][0]) == 1.0 + 3 and float(out["d"][0]) == 10.0 + 2        # d's cohort has seen 2 chunks, a's 3
    assert len(pool) == 4 and "d" in pool and "x" not in pool


def test_leaving_stream_stays_silent_member_until_cohort_is_empty(pool):
    pool.add("a", FakeSession(1.0))
    pool.add("b", FakeSession(2.0))
    pool.step({"a": _chunk(1), "b": _chunk(1)})
    pool.remove("a")
    out = pool.step({"b": _chunk(5)})
    assert set(out) == {"b"} and float(out["b"][0]) == 12.0
    assert pool.cohort_sizes() == [2] and not FakeBatch.log[0].closed             # a is still computed for (silence)
    with pytest.raises(KeyError):
        pool.step({"a": _chunk(1), "b": _chunk(1)})
    pool.remove("b")
    assert pool.n_cohorts == 0 and FakeBatch.log[0].closed and len(pool) == 0


def test_underrun_feeds_silence_and_is_counted(pool):
    pool.add("a", FakeSession(1.0))
    pool.add("b", FakeSession(1.0))
    out = pool.step({"a": _chunk(3)})
    assert float(out["b"][0]) == 1.0 and pool.underruns == {"a": 0, "b": 1}
    with pytest.raises(ValueError):
        pool.step({"a": torch.zeros(100), "b": _chunk(0)})


def test_membership_errors_and_cohort_size_limit():
    FakeBatch.log = []
    pool = StreamPool(decode_chunk_frames=2, max_cohort=2, batch_factory=FakeBatch)
    for i in range(5):
        pool.add(i, FakeSession(1.0))
    with pytest.raises(KeyError):
        pool.add(0, FakeSession(1.0))
    with pytest.raises(ValueError):
        pool.add("no prompt", object())
    pool.remove(4)                                                                 # leaves before it ever ran
    out = pool.step({i: torch.zeros(2, 2048) for i in range(4)})
    assert pool.cohort_sizes() == [2, 2] and out[3].shape == (4096,)
    with pytest.raises(KeyError):
        pool.remove(4)
    pool.close()
    assert all(b.closed for b in FakeBatch.log) and pool.n_cohorts == 0
    with pytest.raises(ValueError):
        StreamPool(max_cohort=0)


class FakeMergeBatch(FakeBatch):
    """A stand-in that can merge: every member keeps its own chunk counter."""

    def __init__(self, sessions):
        super().__init__(sessions)
        self.counts = [0] * len(self.sessions)

    def process_chunk(self, waves):
        assert not self.closed and waves.shape[0] == len(self.sessions)
        self.counts = [c + 1 for c in self.counts]
        return torch.stack([s.gain * w + c for s, w, c in zip(self.sessions, waves, self.counts)])

    @classmethod
    def merged(cls, a, b):
        m = cls(a.sessions + b.sessions)
        m.cfg, m.counts = a.cfg, a.counts + b.counts
        a.close()
        b.close()
        return m


def test_warm_cohorts_merge_and_keep_their_members_state():
    FakeBatch.log = []
    pool = StreamPool(decode_chunk_frames=1, batch_factory=FakeMergeBatch)
    pool.add("a", FakeSession(1.0))
    pool.add("b", FakeSession(2.0))
    pool.step({"a": _chunk(0), "b": _chunk(0)})
    pool.add("c", FakeSession(3.0))                                                # one chunk later: its own cohort
    sizes = []
    for _ in range(6):
        out = pool.step({k: _chunk(0) for k in ("a", "b", "c")})
        sizes.append(pool.cohort_sizes())
    assert sizes[0] == [2, 1] and sizes[-1] == [3] and pool.merges == 1            # merged once c's cohort was warm
    assert float(out["a"][0]) == 7.0 and float(out["c"][0]) == 6.0                 # every member kept its own count
    pool.remove("a")
    out = pool.step({"b": _chunk(1), "c": _chunk(1)})                              # a stays a silent member of the merged cohort
    assert set(out) == {"b", "c"} and pool.cohort_sizes() == [3]
    assert float(out["b"][0]) == 2.0 + 8 and float(out["c"][0]) == 3.0 + 7
    pool.close()


class FakeSelectBatch(FakeMergeBatch):
    """... and can continue with a subset of its members."""

    @classmethod
    def selected(cls, a, keep):
        assert list(keep) == sorted(set(keep)) and all(0 <= k < len(a.sessions) for k in keep)
        m = cls([a.sessions[k] for k in keep])
        m.cfg, m.counts = a.cfg, [a.counts[k] for k in keep]
        a.close()
        return m


def test_silent_members_are_dropped_from_warm_cohorts():
    """Compaction: once the streams that have left are `compact_fraction` of a warm cohort, the cohort continues with its
    live members only (BatchSession.selected); every remaining stream keeps its state and its key -> row mapping."""
    FakeBatch.log = []
    pool = StreamPool(decode_chunk_frames=1, batch_factory=FakeSelectBatch, compact_fraction=0.5)
    sessions = {k: FakeSession(g) for k, g in zip("abcd", (1.0, 2.0, 3.0, 4.0))}
    for k, s in sessions.items():
        pool.add(k, s)
    pool.step({k: _chunk(0) for k in "abcd"})
    pool.remove("b")                                                               # cohort not warm yet: b stays a silent member
    pool.step({k: _chunk(0) for k in "acd"})
    assert pool.cohort_sizes() == [4] and pool.compactions == 0
    for _ in range(3):
        pool.step({k: _chunk(0) for k in "acd"})
    assert pool.cohort_sizes() == [4] and pool.compactions == 0                    # warm, but 1 of 4 is below the fraction
    assert pool.in_use(sessions["b"])
    pool.remove("d")
    out = pool.step({"a": _chunk(1), "c": _chunk(1)})                              # 2 of 4 silent: compacted before this step
    assert pool.cohort_sizes() == [2] and pool.compactions == 1 and len(FakeBatch.log) == 2 and FakeBatch.log[0].closed
    assert float(out["a"][0]) == 1.0 + 6 and float(out["c"][0]) == 3.0 + 6         # counts carried over
    assert not pool.in_use(sessions["b"]) and not pool.in_use(sessions["d"]) and pool.in_use(sessions["c"])
    pool.remove("a")
    out = pool.step({"c": _chunk(2)})                                              # 1 of 2: compacted again, c is row 0 now
    assert pool.cohort_sizes() == [1] and pool.compactions == 2 and float(out["c"][0]) == 6.0 + 7
    pool.remove("c")
    assert pool.n_cohorts == 0 and all(b.closed for b in FakeBatch.log)
    # a factory without `selected` (or compact_fraction = 0) keeps the silent members, as before
    pool = StreamPool(decode_chunk_frames=1, batch_factory=FakeSelectBatch, compact_fraction=0.0)
    for k in "ab":
        pool.add(k, FakeSession(1.0))
    for _ in range(5):
        pool.step({})
    pool.remove("a")
    pool.step({"b": _chunk(0)})
    assert pool.cohort_sizes() == [2] and pool.compactions == 0
    pool.close()

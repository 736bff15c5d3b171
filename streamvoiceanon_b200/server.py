"""Many streams that do not start together ("batched step_batch scheduler with heterogeneous stream
phases").

`BatchSession` advances N streams in lock-step and wants them to start at the same chunk, because the warm-up phases of
the loop (`process_one_chunk`, evaluations/infer_arvc.py:519-525: silent chunks until `delay` ids are there, then the
delay prefill) are shared by the batch.  A server sees streams arrive and leave at any chunk boundary.  `StreamPool`
puts the streams that arrive at the same boundary (and use the same `delay`) into one COHORT = one `BatchSession`;
every `step` advances all cohorts, one library call each.  Once a cohort has left its warm-up chunks it is MERGED into an
older cohort of the same delay (`BatchSession.merged` / svanon_batch_merge: every member's wave ring, encoder and vocoder
state moves into side-by-side buffers), so that in steady state the pool is back to one pass over the weights per chunk
however the streams arrived.  A stream that leaves stays in its cohort as a silent member (the batch-level encoder /
vocoder state is laid out per member) until the cohort is empty, then the cohort is closed -- or until the silent members
are `compact_fraction` of a warm cohort: then the live members continue as a batch of their own
(`BatchSession.selected` / svanon_batch_select), so a long-running server computes for the streams it has.
Each stream produces exactly what it produces alone (the `BatchSession` contract, tests/test_gpu_batch.py).

The reference is one stream per process (`max_batch_size=1`, infer_arvc.py:56); its GUI loop
(real-time-gui.py:1316-1354) is the single-stream case of `step`.
"""
from __future__ import annotations

from typing import Callable, Dict, Hashable, List, Optional

import torch

SAMPLES_PER_FRAME = 2048          # evaluations/infer_arvc.py:28


class _Cohort:
    def __init__(self, keys, sessions, batch, delay):
        self.keys: List[Hashable] = list(keys)
        self.sessions = list(sessions)
        self.batch = batch
        self.delay = delay
        self.live = [True] * len(self.keys)
        self.steps = 0


class StreamPool:
    """pool = StreamPool(encode_window_frames=128, ...)       # the `setup_stream_caches` arguments, shared by all streams
    pool.add(key, session)                                     # a StreamSession whose prompt is set (`set_prompt`)
    out = pool.step({key: chunk, ...})                         # chunk [decode_chunk_frames * 2048] per live stream
    pool.remove(key)

    `step` returns {key: waveform chunk} for every live stream; a live stream without a chunk in this step is fed silence
    and counted in `underruns[key]`.  Streams added since the last step form a new cohort per `delay` value."""

    def __init__(self, encode_window_frames: int = 128, decode_window_frames: int = 64, max_seq_frames: int = 768,
                 buffer_frames: int = 32, decode_chunk_frames: int = 1, max_cohort: int = 256,
                 batch_factory: Optional[Callable] = None, merge_cohorts: bool = True, compact_fraction: float = 0.25):
        if decode_chunk_frames < 1 or max_cohort < 1:
            raise ValueError("decode_chunk_frames and max_cohort must be >= 1")
        self._cfg = dict(encode_window_frames=encode_window_frames, decode_window_frames=decode_window_frames,
                         max_seq_frames=max_seq_frames, buffer_frames=buffer_frames, decode_chunk_frames=decode_chunk_frames)
        self.chunk_samples = decode_chunk_frames * SAMPLES_PER_FRAME
        self.max_cohort = max_cohort
        if batch_factory is None:
            from .streaming import BatchSession
            batch_factory = BatchSession
        self._batch_factory = batch_factory
        self._merge = merge_cohorts and hasattr(batch_factory, "merged")
        self.merges = 0
        self._compact_fraction = float(compact_fraction) if hasattr(batch_factory, "selected") else 0.0
        self.compactions = 0
        self._pending: Dict[Hashable, object] = {}          # key -> session, joined since the last step
        self._cohorts: List[_Cohort] = []
        self._where: Dict[Hashable, tuple] = {}             # key -> (cohort, index)
        self.underruns: Dict[Hashable, int] = {}
        self.steps = 0

    # ------------------------------------------------------------------------------------------ membership
    def add(self, key: Hashable, session) -> None:
        """`session`: a StreamSession after `set_prompt` (its `delay` decides the cohort); it starts with the next `step`."""
        if key in self._pending or key in self._where:
            raise KeyError(f"stream {key!r} is already in the pool")
        if not hasattr(session, "delay"):
            raise ValueError("set the stream's prompt (StreamSession.set_prompt) before adding it to the pool")
        self._pending[key] = session
        self.underruns[key] = 0

    def remove(self, key: Hashable) -> None:
        """The stream stops producing output; its session is NOT closed (the caller owns it) but must stay alive while a
        cohort still advances it as a silent member (`in_use(session)`; False at the latest after `close()`)."""
        if key in self._pending:
            del self._pending[key]
            del self.underruns[key]
            return
        if key not in self._where:
            raise KeyError(f"stream {key!r} is not in the pool")
        cohort, i = self._where.pop(key)
        cohort.live[i] = False
        del self.underruns[key]
        if not any(cohort.live):
            self._retire(cohort)

    def in_use(self, session) -> bool:
        """True while a cohort's batch still holds `session` (as a live or a silent member)."""
        return any(session is s for c in self._cohorts for s in c.sessions) or any(session is s for s in self._pending.values())

    def __contains__(self, key):
        return key in self._pending or key in self._where

    def __len__(self):
        return len(self._pending) + len(self._where)

    @property
    def n_cohorts(self) -> int:
        return len(self._cohorts)

    def cohort_sizes(self) -> List[int]:
        """Members per cohort, silent ones included (what each library call computes for)."""
        return [len(c.keys) for c in self._cohorts]

    def _retire(self, cohort: _Cohort) -> None:
        cohort.batch.close()
        self._cohorts.remove(cohort)

    def _admit(self) -> None:
        by_delay: Dict[int, List[Hashable]] = {}
        for key, sess in self._pending.items():
            by_delay.setdefault(int(sess.delay), []).append(key)
        for delay in sorted(by_delay):
            keys = by_delay[delay]
            for lo in range(0, len(keys), self.max_cohort):
                part = keys[lo: lo + self.max_cohort]
                sessions = [self._pending[k] for k in part]
                batch = self._batch_factory(sessions)
                batch.setup(**self._cfg)
                cohort = _Cohort(part, sessions, batch, delay)
                self._cohorts.append(cohort)
                for i, k in enumerate(part):
                    self._where[k] = (cohort, i)
        self._pending.clear()

    def _warm(self, cohort: _Cohort) -> bool:
        """Past the warm-up chunks of the loop (evaluations/infer_arvc.py:519-525): `delay` ids collected, the delay prefill
        done and one frame decoded -- from then on every chunk of the cohort is a plain decode step."""
        need = 1 if cohort.delay == 0 else -(-cohort.delay // self._cfg["decode_chunk_frames"]) + 1
        return cohort.steps >= need + 1

    def _merge_warm_cohorts(self) -> None:
        """Folds every warm cohort into the oldest warm cohort of the same delay while the result fits max_cohort."""
        by_delay: Dict[int, _Cohort] = {}
        for cohort in list(self._cohorts):
            if not self._warm(cohort):
                continue
            host = by_delay.get(cohort.delay)
            if host is None or len(host.keys) + len(cohort.keys) > self.max_cohort:
                if host is None:
                    by_delay[cohort.delay] = cohort
                continue
            merged = self._batch_factory.merged(host.batch, cohort.batch)
            host.batch = merged
            base = len(host.keys)
            host.keys += cohort.keys
            host.sessions += cohort.sessions
            host.live += cohort.live
            for i, k in enumerate(cohort.keys):
                if k in self._where and self._where[k][0] is cohort:
                    self._where[k] = (host, base + i)
            self._cohorts.remove(cohort)
            self.merges += 1

    def _compact_cohorts(self) -> None:
        """A warm cohort whose silent members have reached `compact_fraction` of its size continues with its live members only."""
        if self._compact_fraction <= 0:
            return
        for cohort in self._cohorts:
            dead = cohort.live.count(False)
            if dead == 0 or dead == len(cohort.live) or dead < self._compact_fraction * len(cohort.live) or not self._warm(cohort):
                continue
            keep = [i for i, live in enumerate(cohort.live) if live]
            cohort.batch = self._batch_factory.selected(cohort.batch, keep)
            cohort.keys = [cohort.keys[i] for i in keep]
            cohort.sessions = [cohort.sessions[i] for i in keep]
            cohort.live = [True] * len(keep)
            for i, k in enumerate(cohort.keys):
                self._where[k] = (cohort, i)
            self.compactions += 1

    # ------------------------------------------------------------------------------------------ the loop
    def step(self, chunks: Dict[Hashable, torch.Tensor]) -> Dict[Hashable, torch.Tensor]:
        """One chunk period for every stream of the pool."""
        for key in chunks:
            if key not in self:
                raise KeyError(f"chunk for unknown stream {key!r}")
        self._compact_cohorts()
        if self._merge:
            self._merge_warm_cohorts()
        self._admit()
        out: Dict[Hashable, torch.Tensor] = {}
        for cohort in list(self._cohorts):
            given = [chunks.get(k) if live else None for k, live in zip(cohort.keys, cohort.live)]
            like = next((g for g in given if g is not None), None)
            device = like.device if like is not None else torch.device("cpu")
            waves = torch.zeros(len(cohort.keys), self.chunk_samples, dtype=torch.float32, device=device)
            for i, (k, g) in enumerate(zip(cohort.keys, given)):
                if g is None:
                    if cohort.live[i]:
                        self.underruns[k] += 1
                    continue
                g = g.reshape(-1)
                if g.numel() != self.chunk_samples:
                    raise ValueError(f"stream {k!r}: chunk of {g.numel()} samples, expected {self.chunk_samples}")
                waves[i] = g.to(device=device, dtype=torch.float32)
            res = cohort.batch.process_chunk(waves)
            cohort.steps += 1
            for i, k in enumerate(cohort.keys):
                if cohort.live[i]:
                    out[k] = res[i]
        self.steps += 1
        return out

    def close(self) -> None:
        for cohort in list(self._cohorts):
            self._retire(cohort)
        self._where.clear()
        self._pending.clear()

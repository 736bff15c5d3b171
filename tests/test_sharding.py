"""N>1 host logic on CPU: two gloo ranks on 127.0.0.1 (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streamvoiceanon_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.streams_for_rank(11, rank, world)
    # each rank "times" a different duration; the job time is the slowest rank's
    dev_ms, e2e_ms = sharding.max_over_ranks([10.0 + 5.0 * rank, 20.0 - 3.0 * rank])
    dist.barrier()
    q.put((rank, mine, dev_ms, e2e_ms))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(s for _, mine, _, _ in out for s in mine)
    assert owned == list(range(11))                      # every stream on exactly one rank
    assert out[0][1] == [0, 2, 4, 6, 8, 10] and out[1][1] == [1, 3, 5, 7, 9]
    for _, _, dev_ms, e2e_ms in out:
        assert dev_ms == 15.0 and e2e_ms == 20.0         # max over ranks, identical everywhere
    assert sharding.aggregate_frames_per_sec(100, 2, 15.0) == 2 * 100 / 0.015


def test_single_process_is_identity():
    assert sharding.max_over_ranks([3.5, 1.0]) == [3.5, 1.0]
    assert sharding.streams_for_rank(5, 0, 1) == [0, 1, 2, 3, 4]

"""world_size-2 gloo test of the N>1 path: contiguous unit sharding + the final token all-gather
(SURVEY.md 8e).  The data path has no collective; this covers the only exchange step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import wb200  # noqa: F401
from whisper_burn_b200 import shard


def test_shard_units_partition():
    for n in (0, 1, 3, 8, 24, 25, 384):
        for g in (1, 2, 4, 8):
            blocks = [shard.shard_units(n, g, r) for r in range(g)]
            covered = [i for s, e in blocks for i in range(s, e)]
            assert covered == list(range(n))
            assert max(e - s for s, e in blocks) == (n + g - 1) // g if n else True


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = shard.shard_units(n_units, world, rank)
        local = [[1000 * u + j for j in range(4 + (u * 7) % 11)] for u in range(s, e)]   # ragged token lists
        out = shard.gather_tokens(local, n_units, capacity=32)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [3, 8])
def test_gather_tokens_gloo_world2(n_units):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[1000 * u + j for j in range(4 + (u * 7) % 11)] for u in range(n_units)]
    assert results[0] == want and results[1] == want

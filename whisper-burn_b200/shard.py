"""Multi-GPU plumbing of the hot path: windows/chunks are independent units
(transcribe.rs:195-203 shadows the cross-window prompt), so they are sharded across ranks in
contiguous blocks with replicated weights and NO data-path collective; the only exchange is one
all-gather of fixed-size token buffers after decoding (NCCL over NVLink on GPUs, gloo in CPU tests).
The reference itself is single-process, single-device (src/bin/transcribe/main.rs:81)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_units(n_units: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of ceil(n_units / world_size) units for `rank` (keeps the
    windows of one chunk on one GPU so the overlap merge stays local)."""
    per = (n_units + world_size - 1) // world_size
    start = min(rank * per, n_units)
    return start, min(start + per, n_units)


def gather_tokens(local: Sequence[Sequence[int]], n_units: int, capacity: int, device: torch.device | str = "cpu",
                  group=None) -> List[List[int]]:
    """All-gathers the per-unit token id lists of every rank (one collective of
    int32[units_per_rank, capacity + 1]; last column = length) and returns them in global unit order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per = (n_units + world - 1) // world
    buf = torch.zeros((per, capacity + 1), dtype=torch.int32)
    for i, toks in enumerate(local):
        if len(toks) > capacity:
            raise ValueError("token list longer than capacity")
        buf[i, :len(toks)] = torch.tensor(list(toks), dtype=torch.int32)
        buf[i, capacity] = len(toks)
    buf = buf.to(device)
    if world == 1:
        gathered = [buf]
    else:
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf, group=group)
    out: List[List[int]] = []
    for r in range(world):
        s, e = shard_units(n_units, world, r)
        g = gathered[r].cpu()
        for i in range(e - s):
            n = int(g[i, capacity])
            out.append([int(v) for v in g[i, :n]])
    return out

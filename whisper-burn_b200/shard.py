"""Multi-GPU plumbing of the hot path: windows/chunks are independent units
(transcribe.rs:195-203 shadows the cross-window prompt), so they are sharded across ranks in
contiguous blocks with replicated weights and NO data-path collective; the only exchange is one
all-gather of fixed-size token buffers after decoding (NCCL over NVLink on GPUs, gloo in CPU tests).
The reference itself is single-process, single-device (src/bin/transcribe/main.rs:81)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_units(n_units: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of ceil(n_units / world_size) units for `rank` (keeps the
    windows of one chunk on one GPU so the overlap merge stays local)."""
    per = (n_units + world_size - 1) // world_size
    start = min(rank * per, n_units)
    return start, min(start + per, n_units)


class TokenGather:
    """The one exchange step, with every buffer allocated once: the rank's token lists are packed into a pinned
    int32[units_per_rank, capacity + 1] array (last column = length), copied to the device in one transfer, all-gathered in ONE
    collective into a single device tensor (ncclAllGather over NVLink; gloo on CPU) and read back with one pinned D2H copy."""

    def __init__(self, n_units: int, capacity: int, device: torch.device | str = "cpu", group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_units, self.capacity = n_units, capacity
        self.per = (n_units + self.world - 1) // self.world
        self.device = torch.device(device)
        cuda = self.device.type == "cuda"
        self.h_in = torch.zeros((self.per, capacity + 1), dtype=torch.int32, pin_memory=cuda)
        self.h_out = torch.zeros((self.world * self.per, capacity + 1), dtype=torch.int32, pin_memory=cuda)
        self.d_in = torch.zeros_like(self.h_in, device=self.device) if cuda else self.h_in
        self.d_out = torch.zeros_like(self.h_out, device=self.device) if cuda else self.h_out
        self.np_in = self.h_in.numpy()
        self.np_out = self.h_out.numpy()

    def __call__(self, local: Sequence[Sequence[int]]) -> np.ndarray:
        """Returns the gathered int32[world * units_per_rank, capacity + 1] array (a view of the pinned buffer)."""
        cap = self.capacity
        self.np_in[:] = 0
        for i, toks in enumerate(local):
            if len(toks) > cap:
                raise ValueError("token list longer than capacity")
            self.np_in[i, :len(toks)] = toks
            self.np_in[i, cap] = len(toks)
        if self.device.type == "cuda":
            self.d_in.copy_(self.h_in, non_blocking=True)
        if self.world == 1:
            self.d_out.copy_(self.d_in)
        else:
            dist.all_gather_into_tensor(self.d_out, self.d_in, group=self.group)
        if self.device.type == "cuda":
            self.h_out.copy_(self.d_out, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return self.np_out

    def as_lists(self, gathered: np.ndarray) -> List[List[int]]:
        """Token lists in global unit order."""
        out: List[List[int]] = []
        for r in range(self.world):
            s, e = shard_units(self.n_units, self.world, r)
            for i in range(e - s):
                row = gathered[r * self.per + i]
                out.append([int(v) for v in row[:int(row[self.capacity])]])
        return out


def gather_tokens(local: Sequence[Sequence[int]], n_units: int, capacity: int, device: torch.device | str = "cpu",
                  group=None) -> List[List[int]]:
    """All-gathers the per-unit token id lists of every rank and returns them in global unit order."""
    g = TokenGather(n_units, capacity, device, group)
    return g.as_lists(g(local))

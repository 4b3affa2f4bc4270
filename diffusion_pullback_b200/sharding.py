"""Multi-GPU plumbing for the pullback path: the (x_t, t, prompt) problems are independent (the reference loops over
them in separate processes, `src/scripts/*.sh:1-6`, `src/main.py:61-76`), so they are dealt round-robin to the ranks
(one process per GPU, full weight replica each, no data-path collective) and ONE all-gather at the end hands every rank
the singular values / right singular vectors of every problem.  Backend: NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_problems(n_problems: int, rank: int, world: int):
    """Indices of the problems rank `rank` solves (round-robin keeps ragged counts within one of each other)."""
    return list(range(rank, n_problems, world))


def gather_results(local, n_problems: int, k: int, n_in: int, device, group=None):
    """local: [(problem index, s [k], vT [k, n_in])] solved on this rank -> {index: (s, vT)} for ALL problems on every
    rank, through a single all_gather of one padded [slots, 1 + k + k*n_in] fp32 payload per rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    slots = (n_problems + world - 1) // world
    width = 1 + k + k * n_in
    payload = torch.full((slots, width), -1.0, dtype=torch.float32, device=device)
    for j, (idx, s, vT) in enumerate(local):
        payload[j, 0] = float(idx)
        payload[j, 1:1 + k] = s.reshape(-1)
        payload[j, 1 + k:] = vT.reshape(-1)
    if world > 1:
        bufs = [torch.empty_like(payload) for _ in range(world)]
        dist.all_gather(bufs, payload, group=group)
    else:
        bufs = [payload]
    out = {}
    for b in bufs:
        for row in b:
            idx = int(row[0].item())
            if idx >= 0:
                out[idx] = (row[1:1 + k].clone(), row[1 + k:].reshape(k, n_in).clone())
    return out

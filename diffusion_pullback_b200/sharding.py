"""Multi-GPU plumbing for the pullback path: the (x_t, t, prompt) problems are independent (the reference loops over
them in separate processes, `src/scripts/*.sh:1-6`, `src/main.py:61-76`), so they are dealt round-robin to the ranks
(one process per GPU, full weight replica each, no data-path collective) and ONE all-gather at the end hands every rank
the singular values / right singular vectors of every problem.  Backend: NCCL on GPUs, gloo in the CPU tests.

When there are fewer problems than GPUs (BASELINE.json configs[3]: 10 timesteps on 8 GPUs; configs[4]: one prompt) the
k tangent columns of ONE problem are split over a group of ranks as well (`plan_2d`, `pullback_tangent_sharded`): the
columns are independent inside an iteration (U_g = J V[:, g], W_g = U_g^T J), so each rank pushes only its own columns
through the U-Net and one small all-gather of the W rows (k x n_in fp32, <= 3.3 MB) per iteration precedes the
redundant k x n_in re-orthonormalisation on every rank."""
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


def plan_2d(n_problems: int, world: int):
    """(tangent-group size g, number of groups): the smallest g dividing `world` such that the world/g groups split the
    problems evenly.  10 problems on 8 GPUs -> g = 4 (2 groups x 5 problems, k/4 columns per GPU); 16 on 8 -> g = 1."""
    import math
    if n_problems <= 0:
        return 1, world
    g = world // math.gcd(n_problems, world)
    return g, world // g


def shard_columns(k: int, rank: int, size: int):
    """[lo, hi) of the tangent columns rank `rank` of a `size`-rank group owns (contiguous, ragged within one)."""
    base, extra = divmod(k, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _all_gather_rows(local_rows, k: int, width: int, group, device):
    """All-gather ragged row blocks ([hi-lo, width] per rank, contiguous column ownership) into [k, width]."""
    size = dist.get_world_size(group)
    slots = (k + size - 1) // size
    pad = torch.zeros(slots, width, dtype=torch.float32, device=device)
    pad[:local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.empty(k, width, dtype=torch.float32, device=device)
    for r in range(size):
        lo, hi = shard_columns(k, r, size)
        out[lo:hi] = bufs[r][:hi - lo]
    return out


@torch.no_grad()
def pullback_tangent_sharded(eng, V0, min_iter: int, max_iter: int, tol: float, group=None):
    """`utils.py:756-808` for ONE problem with its k tangent columns split over the ranks of `group`.  Every rank must
    have called `eng.set_point` with the same (x_t, t, prompt) (the 261 GF primal pass is recomputed per rank, SURVEY.md
    s.8e) and passes the same V0 [k, n_in].  Returns (u [k, n_out], s [k], vT [k, n_in], info) on every rank."""
    size = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    V0 = V0.to(dtype=torch.float32).reshape(-1, eng.n_in).contiguous()
    k = V0.shape[0]
    if size == 1:
        return eng.pullback(V0, min_iter, max_iter, tol)
    lo, hi = shard_columns(k, rank, size)
    dev = V0.device
    V, s, Ul = V0, None, None
    done, converged, last = 0, False, 0.0
    for i in range(max_iter):
        if hi > lo:
            Ul = eng.jvp(V[lo:hi])                            # [hi-lo, n_out]
            Wl = eng.vjp(Ul)                                  # [hi-lo, n_in]
        else:
            Ul = torch.empty(0, eng.n_out, device=dev)
            Wl = torch.empty(0, eng.n_in, device=dev)
        W = _all_gather_rows(Wl, k, eng.n_in, group, dev)     # the per-iteration exchange
        s, Vn, met = eng.orthonormalize(W, V, tol)            # redundant on every rank
        done += 1
        last_it = i + 1 == max_iter
        if i > min_iter or last_it:                           # the reference's early-exit test (utils.py:806-808)
            m = met.clone()
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)   # one decision for the whole group
            last = float(m[0]) ** 0.5
            converged = i > min_iter and float(m[1]) == 0.0
        V = Vn
        if converged or last_it:
            break
    u = _all_gather_rows(Ul, k, eng.n_out, group, dev)        # end of the problem: left vectors of every column
    from . import _native as N
    info = N.PbIterInfo()
    info.iters_done, info.converged, info.last_dist = done, int(converged), last
    return u, s, V, info


def plan_rounds(n_problems: int, world: int):
    """Mixed partitioning of P problems over N ranks (both partitionings in one schedule): full rounds of one
    problem per rank, then the P mod N left-over problems each split over N // (P mod N) ranks by tangent columns, so that no
    GPU idles through a whole round.  Returns a list of rounds, each a list of (problem index, tuple of ranks).
    10 problems on 8 GPUs: [[(0, (0,)), ..., (7, (7,))], [(8, (0, 1, 2, 3)), (9, (4, 5, 6, 7))]]."""
    rounds, p = [], 0
    while n_problems - p >= world:
        rounds.append([(p + r, (r,)) for r in range(world)])
        p += world
    left = n_problems - p
    if left > 0:
        g = world // left
        rounds.append([(p + j, tuple(range(j * g, (j + 1) * g))) for j in range(left)])
    return rounds


def make_round_groups(plan):
    """torch.distributed groups of the multi-rank entries of a `plan_rounds` schedule ({ranks: group}); collective: every rank
    calls it with the same plan."""
    groups = {}
    for rnd in plan:
        for _, ranks in rnd:
            if len(ranks) > 1 and ranks not in groups:
                groups[ranks] = dist.new_group(list(ranks))
    return groups


@torch.no_grad()
def solve_rounds(eng, plan, groups, rank: int, set_point, v0_of, min_iter: int, max_iter: int, tol: float):
    """Run a `plan_rounds` schedule on this rank: `set_point(i)` sets problem i on `eng`, `v0_of(i)` is its start subspace
    [k, n_in] (identical on every rank of a group).  Returns [(problem index, s, vT)] for the problems this rank solved or
    co-solved as the FIRST rank of its group (so that a later `gather_results` sees each problem once)."""
    out = []
    for rnd in plan:
        for idx, ranks in rnd:
            if rank not in ranks:
                continue
            set_point(idx)
            if len(ranks) == 1:
                u, s, vT, info = eng.pullback(v0_of(idx), min_iter, max_iter, tol)
                out.append((idx, s, vT))
            else:
                u, s, vT, info = pullback_tangent_sharded(eng, v0_of(idx), min_iter, max_iter, tol, groups[ranks])
                if rank == ranks[0]:
                    out.append((idx, s, vT))
    return out

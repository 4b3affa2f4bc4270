"""World-size-2 gloo test (CPU) of the N > 1 host logic: problems dealt round-robin, one all-gather of (s, vT)."""
import ctypes as C
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_pullback_b200.sharding import (gather_results, plan_2d, pullback_tangent_sharded, shard_columns,
                                              shard_problems)


def _solve(problem, L):
    from diffusion_pullback_b200.engine import PullbackEngine, unet_config
    from oracle import unet_torch as UT
    m = UT.build_unet("uncond_tiny")
    x, t, _ = UT.synthetic_inputs("uncond_tiny", seed=1234 + problem)
    eng = PullbackEngine(unet_config(m), 32, 32, "mid", 0, 2, 0, "cpu", _lib=L)
    eng.bind(m.state_dict())
    eng.set_point(x, float(t), None)
    g = torch.Generator().manual_seed(problem)
    q, _ = torch.linalg.qr(torch.randn(eng.n_in, 2, generator=g))
    u, s, vT, _ = eng.pullback(q.T.contiguous(), 2, 2, 0.0)
    return s, vT


def _worker(rank, world, port, n_problems, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.hostsim.build import build
    L = C.CDLL(build())
    mine = shard_problems(n_problems, rank, world)
    local = [(p,) + _solve(p, L) for p in mine]
    allr = gather_results(local, n_problems, 2, 3 * 32 * 32, "cpu")
    if rank == 0:
        q.put({k: (v[0], v[1]) for k, v in allr.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_problems_covers_everything_once():
    for n in (0, 1, 3, 8, 10):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in shard_problems(n, r, w))
            assert got == list(range(n))
            sizes = [len(shard_problems(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gather_all_problems():
    from tests.hostsim.build import build
    build()
    n_problems, world = 3, 2                                  # ragged: rank 0 solves 2, rank 1 solves 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_problems, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [0, 1, 2]
    L = C.CDLL(build())
    for p in range(n_problems):
        s, vT = _solve(p, L)
        assert torch.allclose(res[p][0], s, rtol=1e-5) and torch.allclose(res[p][1].abs(), vT.abs(), atol=1e-5)


# ---- tangent sharding inside one problem (secondary partitioning) ----
def test_plan_2d_and_column_shards():
    assert plan_2d(10, 8) == (4, 2)          # BASELINE configs[3]: 10 timesteps on 8 GPUs -> 2 groups of 4
    assert plan_2d(10, 4) == (2, 2) and plan_2d(10, 2) == (1, 2) and plan_2d(1, 8) == (8, 1) and plan_2d(16, 8) == (1, 8)
    for k in (1, 2, 5, 16):
        for size in (1, 2, 3, 4, 8):
            spans = [shard_columns(k, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == k
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _tangent_setup(L, k):
    from diffusion_pullback_b200.engine import PullbackEngine, unet_config
    from oracle import unet_torch as UT
    m = UT.build_unet("uncond_tiny")
    x, t, _ = UT.synthetic_inputs("uncond_tiny", seed=77)
    eng = PullbackEngine(unet_config(m), 32, 32, "mid", 0, k, 0, "cpu", _lib=L)
    eng.bind(m.state_dict())
    eng.set_point(x, float(t), None)
    g = torch.Generator().manual_seed(5)
    q, _ = torch.linalg.qr(torch.randn(eng.n_in, k, generator=g))
    return eng, q.T.contiguous()


def _tangent_worker(rank, world, port, k, iters, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.hostsim.build import build
    eng, v0 = _tangent_setup(C.CDLL(build()), k)
    u, s, vT, info = pullback_tangent_sharded(eng, v0, iters, iters, 0.0, group=None)
    q.put((rank, u, s, vT, info.iters_done))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_split_the_columns_of_one_problem():
    from tests.hostsim.build import build
    L = C.CDLL(build())
    k, iters, world = 3, 3, 2                                 # ragged: rank 0 owns 2 columns, rank 1 owns 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 500
    procs = [ctx.Process(target=_tangent_worker, args=(r, world, port, k, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eng, v0 = _tangent_setup(L, k)
    u, s, vT, info = eng.pullback(v0, iters, iters, 0.0)      # the unsharded loop (pb_pullback)
    for _, u_r, s_r, vT_r, done in got:
        assert done == iters
        assert torch.allclose(s_r, s, rtol=1e-5)
        assert torch.allclose(vT_r, vT, atol=1e-5) and torch.allclose(u_r, u, rtol=1e-4, atol=1e-5)


# ---- mixed schedule: full rounds of one problem per rank, the left-over problems split by tangent columns ----
def test_plan_rounds():
    from diffusion_pullback_b200.sharding import plan_rounds
    p = plan_rounds(10, 8)                    # BASELINE configs[3]: 10 timesteps on 8 GPUs
    assert p[0] == [(i, (i,)) for i in range(8)] and p[1] == [(8, (0, 1, 2, 3)), (9, (4, 5, 6, 7))]
    assert plan_rounds(10, 4) == [[(i, (i,)) for i in range(4)], [(4 + i, (i,)) for i in range(4)], [(8, (0, 1)), (9, (2, 3))]]
    assert plan_rounds(1, 8) == [[(0, tuple(range(8)))]] and plan_rounds(10, 1) == [[(i, (0,))] for i in range(10)]
    assert plan_rounds(3, 8) == [[(0, (0, 1)), (1, (2, 3)), (2, (4, 5))]]
    for n in (0, 1, 3, 7, 10, 16):
        for w in (1, 2, 4, 8):
            seen = sorted(i for rnd in plan_rounds(n, w) for i, _ in rnd)
            assert seen == list(range(n))
            for rnd in plan_rounds(n, w):
                used = [r for _, ranks in rnd for r in ranks]
                assert len(used) == len(set(used)) and all(0 <= r < w for r in used)


def _mixed_worker(rank, world, port, n_problems, k, iters, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_pullback_b200.engine import PullbackEngine, unet_config
    from diffusion_pullback_b200.sharding import make_round_groups, plan_rounds, solve_rounds
    from oracle import unet_torch as UT
    from tests.hostsim.build import build
    L = C.CDLL(build())
    m = UT.build_unet("uncond_tiny")
    eng = PullbackEngine(unet_config(m), 32, 32, "mid", 0, k, 0, "cpu", _lib=L)
    eng.bind(m.state_dict())
    plan = plan_rounds(n_problems, world)
    groups = make_round_groups(plan)

    def set_point(i):
        x, t, _ = UT.synthetic_inputs("uncond_tiny", seed=1234 + i)
        eng.set_point(x, float(t), None)

    def v0_of(i):
        qq, _ = torch.linalg.qr(torch.randn(eng.n_in, k, generator=torch.Generator().manual_seed(i)))
        return qq.T.contiguous()

    local = solve_rounds(eng, plan, groups, rank, set_point, v0_of, iters, iters, 0.0)
    allr = gather_results(local, n_problems, k, eng.n_in, "cpu")
    if rank == 0:
        q.put({kk: (v[0], v[1]) for kk, v in allr.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_mixed_schedule_matches_one_rank():
    """3 problems on 2 ranks: one round of one problem per rank, then problem 2 with its 3 columns split 2 + 1 over both ranks;
    every problem's (s, vT) equals the unsharded pb_pullback run."""
    from diffusion_pullback_b200.engine import PullbackEngine, unet_config
    from oracle import unet_torch as UT
    from tests.hostsim.build import build
    L = C.CDLL(build())
    n_problems, world, k, iters = 3, 2, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 13) % 500
    procs = [ctx.Process(target=_mixed_worker, args=(r, world, port, n_problems, k, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [0, 1, 2]
    m = UT.build_unet("uncond_tiny")
    eng = PullbackEngine(unet_config(m), 32, 32, "mid", 0, k, 0, "cpu", _lib=L)
    eng.bind(m.state_dict())
    for i in range(n_problems):
        x, t, _ = UT.synthetic_inputs("uncond_tiny", seed=1234 + i)
        eng.set_point(x, float(t), None)
        qq, _ = torch.linalg.qr(torch.randn(eng.n_in, k, generator=torch.Generator().manual_seed(i)))
        u, s, vT, _ = eng.pullback(qq.T.contiguous(), iters, iters, 0.0)
        assert torch.allclose(res[i][0], s, rtol=1e-5) and torch.allclose(res[i][1].abs(), vT.abs(), atol=1e-5), i

#!/usr/bin/env python
"""bench.py -- pullback JVP-iters/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic problem: the primal pass at (x_t, t, prompt) plus the full
subspace iteration of `local_encoder_pullback_zt` (BASELINE.json configs[1]: SD-v1.5 512^2, mid-block_0, pca_rank 5,
50 power iterations, edit_t = 0.7 T; random-init weights, synthetic 4x64x64 latent).  One JVP-iter = k JVP columns +
k VJP columns + re-orthonormalisation; value = (ranks x steps x iterations) / seconds.

  value     : inputs (x_t, ctx, V0) resident in HBM before the timed region, device-timed with CUDA events, max over ranks
  e2e       : the same steps through the C-ABI host entry pb_pullback_host (pinned HOST buffers in and out)
  roofline  : the dominant kernel (gemm_tc_kernel, the tcgen05 implicit-GEMM behind every conv / linear / attention product):
              algorithmic flops (2 M N K per product) / device time of its launches, measured live with an event pair
              around every launch of two eagerly launched iterations (pb_profile_*), against the measured dense bf16 peak of
              MEASURED_PEAKS.json (the MMA kind used is TF32, nominally half that peak); `traffic` = DRAM bytes per launch from
              the ncu pass committed under profiles/; `step_*` = the whole step (2 k F_tan per iteration)
  cpu_baseline (N = 1, rank 0): the oracle port of the reference algorithm on torch-CPU, one iteration
  --impl reference: the reference's CPU path (oracle port) alone, on the same config / metric

Multi-GPU: independent problems are sharded over ranks (weak scaling, no data-path collective); one NCCL all-gather of the
singular values / vectors of every solved problem closes the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Libraries write banners there (NCCL prints its version line on the first
# collective whatever NCCL_DEBUG_FILE says), so the process-level stdout is pointed at stderr and the JSON line goes out
# through a saved copy of the original descriptor.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


# name: (model config, op, block_idx, k, iterations, F_tan GF, primal GF)
WORKLOADS = {
    "sd15_mid_k5_i50": ("sd15", "mid", 0, 5, 50, 309.1, 261.4),
    "sd15_up1_k5_i50": ("sd15", "up", 1, 5, 50, 486.5, 438.7),
    "sd15_mid_k16_i50": ("sd15", "mid", 0, 16, 50, 309.1, 261.4),
    "sd21_768_mid_k5_i50": ("sd21_768", "mid", 0, 5, 50, 971.1, 724.8),
    "celebahq_mid_k2_i10": ("celebahq", "mid", 0, 2, 10, 135.3, 135.2),
    "sd_small_mid_k5_i50": ("sd_small", "mid", 0, 5, 50, 0.0, 0.0),
}
METRIC, UNIT = "pullback JVP-iters/sec", "iters/s"


def peaks():
    """Roofline denominators: 16-bit tensor peak as burst (a kernel timed in isolation) and sustained (a kernel inside a long
    step) from MEASURED_PEAKS.json (driver-written), HBM GB/s, and the TF32 peak measured with the same recipe by
    scripts/measure_peaks.py (profiles/measured_tf32_peak.json); without that file TF32 counts as half the 16-bit rate."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        out = {"burst": d["bf16_tflops"], "sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "hbm": d["hbm_gbs"],
               "source": "measured (MEASURED_PEAKS.json: bf16 burst for per-launch fractions, sustained for the whole step)"}
    else:
        out = {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    t = os.path.join(ROOT, "profiles", "measured_tf32_peak.json")
    if os.path.exists(t):
        d = json.load(open(t))
        out.update(tf32_burst=d["tf32_tflops"], tf32_sustained=d["tf32_tflops_sustained"],
                   tf32_source="measured (profiles/measured_tf32_peak.json, scripts/measure_peaks.py)")
    else:
        out.update(tf32_burst=out["burst"] / 2, tf32_sustained=out["sustained"] / 2, tf32_source="nominal (half the 16-bit rate)")
    return out


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [c.strip() for c in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_iteration(model_name, op, bi, k_s, iters=1, threads=None):
    """ORACLE LEG (test infrastructure used as the CPU baseline): the reference algorithm (utils.py:722-816 restated in
    oracle/pullback_oracle.py) on torch-CPU fp32 with all host threads.  Returns seconds per subspace iteration at rank k_s."""
    from oracle import pullback_oracle as PO
    from oracle import unet_torch as UT
    torch.set_num_threads(threads or os.cpu_count())
    m = UT.build_unet(model_name, build_up=(op == "up"))
    x, t, ctx = UT.synthetic_inputs(model_name)
    torch.manual_seed(0)
    v0 = PO.initial_subspace(x.numel(), k_s)
    t0 = time.perf_counter()
    PO.local_encoder_pullback(m, x, t, ctx, op, bi, k_s, iters, iters, 0.0, v0=v0)
    return (time.perf_counter() - t0) / iters


def gpu_reference_iteration(model_name, op, bi, k, tf32):
    """ORACLE LEG: SURVEY.md s.8d(ii), "the reference on B200" -- the oracle port of the reference algorithm through torch eager
    autograd on cuda:0 (dual-number forward over the k tangents + k backward passes, cuDNN / cuBLAS), fp32 with TF32 off (the
    reference's torch-2.1 default) or on (the operand precision of this repo's own path).  Returns seconds per subspace
    iteration at rank k (one warm-up iteration at rank 1, one timed at rank k)."""
    from oracle import pullback_oracle as PO
    from oracle import unet_torch as UT
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = bool(tf32)
    try:
        m = UT.build_unet(model_name, build_up=(op == "up")).to("cuda:0")
        x, t, ctx = UT.synthetic_inputs(model_name)
        x, t = x.to("cuda:0"), t.to("cuda:0")
        ctx = None if ctx is None else ctx.to("cuda:0")
        dt = None
        for k_s in (1, k):
            torch.manual_seed(0)
            v0 = PO.initial_subspace(x.numel(), k_s, device="cuda:0")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            PO.local_encoder_pullback(m, x, t, ctx, op, bi, k_s, 1, 1, 0.0, v0=v0)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        return dt
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        del m
        torch.cuda.empty_cache()


def run_reference(args, wl):
    model_name, op, bi, k, iters, f_tan, f_primal = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pullback_oracle as PO
    from oracle import unet_torch as UT
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    m = UT.build_unet(model_name, build_up=(op == "up"))
    x, t, ctx = UT.synthetic_inputs(model_name)
    on_gpu = args.ref_device == "cuda"
    if on_gpu:
        # the same reference algorithm through torch eager autograd on this box's GPU (fp32, TF32 off like
        # torch 2.1's matmul default) -- "the reference on B200"; not the driver's reference arm (that one is --ref-device cpu)
        tf32 = os.environ.get("PB_REF_TF32", "0") == "1"     # 1: let cuBLAS / cuDNN use TF32 (the operand precision of our own path)
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        m = m.to("cuda:0")
        x, t = x.to("cuda:0"), t.to("cuda:0")
        ctx = None if ctx is None else ctx.to("cuda:0")

    def iteration(k_s, seed):
        torch.manual_seed(seed)
        v0 = PO.initial_subspace(x.numel(), k_s)
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        PO.local_encoder_pullback(m, x, t, ctx, op, bi, k_s, 1, 1, 0.0, v0=v0)
        if on_gpu:
            torch.cuda.synchronize()
        return time.perf_counter() - t0

    t_col = iteration(1, 0)                                  # probe (also warms the thread pool)
    budget = 200.0
    k_s = 1
    for cand in (k, 2, 1):
        if cand <= k and (args.steps + args.warmup) * cand * t_col <= budget:
            k_s = cand
            break
    for i in range(args.warmup):
        iteration(k_s, 100 + i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        iteration(k_s, 200 + i)
    dt = (time.perf_counter() - t0) / args.steps
    per_iter = dt * (k / k_s)                                # reference cost per iteration is linear in k (k dual forwards + k backwards)
    value = 1.0 / per_iter
    sample = (f"each step = ONE subspace iteration (1 of {iters}) at rank {k_s} of {k} on the full {model_name} {op}-{bi} problem"
              + ("" if k_s == k else f", scaled x{k}/{k_s} to rank {k}"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "model": model_name, "op": op, "block_idx": bi, "pca_rank": k, "power_iters": iters,
                       "device": ("cuda:0 (torch eager autograd fp32, TF32 off, oracle port of the reference algorithm)" if on_gpu else
                                  "host CPU (torch-cpu autograd, oracle port of the reference algorithm)")},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if on_gpu:
        line["impl"] = "reference_gpu"
        line["reference_gpu"] = dict(line.pop("cpu_baseline"), device=torch.cuda.get_device_name(0), tf32=tf32,
                                     peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
        line["config"]["device"] = line["config"]["device"].replace("TF32 off", "TF32 on" if tf32 else "TF32 off")
        del line["reference_gpu"]["cores"], line["reference_gpu"]["kind"]
    emit(line)


def run_secondary(dev, rank, world):
    """BASELINE.json configs[3] -- SD-v1.5 mid-block, pca_rank 16, 50 power iterations, 10 edit timesteps -- on the `world` GPUs of
    this job through the MIXED schedule of sharding.plan_rounds: full rounds of one problem per GPU, the left-over problems split
    over groups of GPUs by tangent columns (SURVEY.md s.8e secondary partitioning; one all-gather of the W rows per iteration
    inside a group), one closing all-gather of (s, vT).  With one GPU: the ten problems one after the other -- the denominator
    of the scaling the driver computes from the per-N records.  Device-timed, max over ranks."""
    import torch.distributed as dist
    import diffusion_pullback_b200 as PB
    from diffusion_pullback_b200 import synthetic as SY
    from diffusion_pullback_b200.sharding import gather_results, make_round_groups, plan_rounds, solve_rounds
    model_name, op, bi, k, iters, nprob = "sd15", "mid", 0, 16, 50, 10
    unet = SY.SyntheticUNet(model_name, upto=(op, bi), device=dev)
    cfg = PB.unet_config(unet)
    size, ctx_len = unet.config["sample_size"], unet.config["ctx_len"]
    eng = PB.PullbackEngine(cfg, size, size, op, bi, k, ctx_len, dev)
    eng.bind(unet.state_dict())
    unet._sd = None
    torch.cuda.empty_cache()
    _, _, ctx = SY.synthetic_inputs(model_name)
    ctxd = ctx.to(dev)
    # ten edit timesteps of one image (0.95 T ... 0.5 T) with the latent each of them is reached at (synthetic)
    ts = [999.0 * (0.95 - 0.05 * i) for i in range(nprob)]
    xs = [torch.randn(1, cfg["in_channels"], size, size, generator=torch.Generator().manual_seed(4321 + i)).to(dev) for i in range(nprob)]
    v0 = {}

    def v0_of(i):
        if i not in v0:
            q, _ = torch.linalg.qr(torch.randn(eng.n_in, k, generator=torch.Generator().manual_seed(99 + i)))
            v0[i] = q.T.contiguous().to(dev)
        return v0[i]

    plan = plan_rounds(nprob, world)
    groups = make_round_groups(plan) if world > 1 else {}
    for i in range(nprob):
        v0_of(i)
    # warm-up: every code path this rank will take (graph capture of the iteration, of the stand-alone jvp / vjp at the group's
    # column count), two iterations each
    warm = [[(idx, ranks) for idx, ranks in rnd if rank in ranks][:1] for rnd in plan]
    seen = set()
    for rnd in warm:
        for idx, ranks in rnd:
            if len(ranks) in seen:
                continue
            seen.add(len(ranks))
            solve_rounds(eng, [[(idx, ranks)]], groups, rank, lambda i: eng.set_point(xs[i], ts[i], ctxd), v0_of, 2, 2, 0.0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    local = solve_rounds(eng, plan, groups, rank, lambda i: eng.set_point(xs[i], ts[i], ctxd), v0_of, iters, iters, 0.0)
    allr = gather_results(local, nprob, k, eng.n_in, dev) if world > 1 else {i: (s_, v_) for i, s_, v_ in local}
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm)
    assert sorted(allr) == list(range(nprob))
    del eng
    torch.cuda.empty_cache()
    return {"workload": "sd15_mid_k16_i50 x 10 edit timesteps (BASELINE.json configs[3])", "value": nprob * iters / (ms / 1e3), "unit": UNIT,
            "seconds": ms / 1e3, "n_gpus": world, "problems": nprob, "pca_rank": k, "power_iters": iters,
            "schedule": [[[idx, list(ranks)] for idx, ranks in rnd] for rnd in plan],
            "collectives": "one all-gather of the W rows (k x n_in fp32 = 1.0 MB) per iteration inside a tangent group (NCCL), "
                           "one all-gather of (s, vT) of the ten problems at the end",
            "sigma_max_first_problem": float(allr[0][0][0])}


def run_ours(args, wl):
    import torch.distributed as dist
    import diffusion_pullback_b200 as PB
    from diffusion_pullback_b200 import synthetic as SY
    model_name, op, bi, k, iters, f_tan, f_primal = WORKLOADS[wl]
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, args.warmup
    unet = SY.SyntheticUNet(model_name, upto=(op, bi), device=dev)
    cfg = PB.unet_config(unet)
    size, ctx_len = unet.config["sample_size"], unet.config["ctx_len"]
    # problems per step, batched through pb_set_slots when > 1.  Default: as many queued problems as fit one handle (k_max = 64
    # columns), at most 5 -- the metric counts problems x iterations per second, the reference's own driver loops
    # over independent problems (main.py:71-76), and one problem at a time is reported beside it (`slots1`)
    S = max(1, args.slots) if args.slots else max(1, min(5, 64 // k))
    if args.shard == "tangent":
        S = 1
    if S > 1 and (args.shard == "tangent" or S * k > 64):
        raise SystemExit("bench.py: --slots needs problem sharding and slots * pca_rank <= 64")
    eng = PB.PullbackEngine(cfg, size, size, op, bi, k, ctx_len, dev)
    eng.bind(unet.state_dict())
    engS = None
    while S > 1:                                             # as many slots as fit the HBM (SD-2.1-768: 12.4 GB of primal cache per slot)
        try:
            engS = PB.PullbackEngine(cfg, size, size, op, bi, S * k, ctx_len, dev)
            engS.bind(unet.state_dict())
            engS.set_slots(S)
            if torch.cuda.mem_get_info(dev)[0] < 8 << 30:   # head-room for the e2e buffers and the probe pass
                raise torch.OutOfMemoryError("not enough head-room")
            break
        except torch.OutOfMemoryError:
            if args.slots:
                raise SystemExit(f"bench.py: --slots {S} does not fit the device memory")
            engS = None
            torch.cuda.empty_cache()
            S -= 1
    unet._sd = None                                          # packed copy lives in the engine now
    torch.cuda.empty_cache()
    _, t, ctx = SY.synthetic_inputs(model_name)
    tval = float(t)
    # one independent problem per (rank, step): synthetic x_t and the reference's randn + QR start (utils.py:750-752)
    tangent = args.shard == "tangent" and world > 1          # every rank works on the SAME problem and owns k / world of its columns
    prank = 0 if tangent else rank
    xs, v0s = [], []
    for i in range((K + Wm) * S):
        g = torch.Generator().manual_seed(1234 + 1000 * i + prank)
        xs.append(torch.randn(1, cfg["in_channels"], size, size, generator=g))
        g2 = torch.Generator().manual_seed(i * world + prank)
        q, _ = torch.linalg.qr(torch.randn(eng.n_in, k, generator=g2))
        v0s.append(q.T.contiguous())
    xd = [x.to(dev) for x in xs]
    v0d = [v.to(dev) for v in v0s]
    ctxd = ctx.to(dev) if ctx is not None else None
    results = []

    def step(i):
        if S > 1:                                            # S problems, one tangent batch of S * k columns
            for p in range(S):
                engS.set_point(xd[i * S + p], tval, ctxd, slot=p)
            u, s, vT, info = engS.pullback(torch.cat(v0d[i * S:(i + 1) * S], 0), iters, iters, 0.0)
            return s, vT
        eng.set_point(xd[i], tval, ctxd)
        if tangent:
            from diffusion_pullback_b200.sharding import pullback_tangent_sharded
            u, s, vT, info = pullback_tangent_sharded(eng, v0d[i], iters, iters, 0.0)
        else:
            u, s, vT, info = eng.pullback(v0d[i], iters, iters, 0.0)
        return s, vT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = engS.launches if S > 1 else eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(Wm, Wm + K):
        results.append(step(i))
    payload = torch.cat([torch.cat([s, vT.reshape(-1)]) for s, vT in results])
    if world > 1 and not tangent:
        gathered = [torch.empty_like(payload) for _ in range(world)]
        dist.all_gather(gathered, payload)                   # the single collective: singular values + vectors of every problem
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = (engS.launches if S > 1 else eng.launches) - l0
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax)
    secs = ms / 1e3
    nprob = K if tangent else world * K * S                   # problems solved by the whole job
    value = nprob * iters / secs

    # ---- one problem at a time (the reference's granularity) beside the batched number ----
    slots1 = None
    if S > 1:
        n1 = max(2, min(K, 4))
        for i in range(2):
            eng.set_point(xd[i], tval, ctxd); eng.pullback(v0d[i], iters, iters, 0.0)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(n1):
            eng.set_point(xd[i], tval, ctxd)
            eng.pullback(v0d[i], iters, iters, 0.0)
        f1.record()
        barrier()
        ms1 = f0.elapsed_time(f1)
        if world > 1:
            tm = torch.tensor([ms1], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms1 = float(tm)
        slots1 = {"value": world * n1 * iters / (ms1 / 1e3), "unit": UNIT, "steps": n1, "ms_per_problem": ms1 / n1,
                  "note": "one problem per pb_pullback call on the same GPU(s), device-timed"}

    # ---- e2e: the C-ABI host entry with pinned host buffers, H2D / D2H inside the timed region ----
    xh = [x.contiguous().pin_memory() for x in xs]
    v0h = [v.pin_memory() for v in v0s]
    ctxh = ctx.contiguous().pin_memory() if ctx is not None else None
    out = (torch.empty(k, eng.n_out).pin_memory(), torch.empty(k).pin_memory(), torch.empty(k, eng.n_in).pin_memory())
    if S > 1:                                                # host buffers of S problems per step (pb_pullback_host_slots)
        xhS = [torch.cat([xs[i * S + p].reshape(1, -1) for p in range(S)], 0).contiguous().pin_memory() for i in range(K + Wm)]
        v0hS = [torch.cat(v0s[i * S:(i + 1) * S], 0).contiguous().pin_memory() for i in range(K + Wm)]
        ctxhS = ctx.repeat(S, 1, 1).contiguous().pin_memory() if ctx is not None else None
        outS = (torch.empty(S * k, eng.n_out).pin_memory(), torch.empty(S * k).pin_memory(), torch.empty(S * k, eng.n_in).pin_memory())

    def e2e_step(i):
        if S > 1:
            engS.pullback_host(xhS[i], [tval] * S, ctxhS, v0hS[i], iters, iters, 0.0, out=outS)
            return
        if not tangent:
            eng.pullback_host(xh[i], tval, ctxh, v0h[i], iters, iters, 0.0, out=out)
            return
        # tangent-sharded: host buffers in, the public sharded call, host buffers out
        from diffusion_pullback_b200.sharding import pullback_tangent_sharded
        x = xh[i].to(dev, non_blocking=True)
        v0 = v0h[i].to(dev, non_blocking=True)
        c = ctxh.to(dev, non_blocking=True) if ctxh is not None else None
        eng.set_point(x, tval, c)
        u, s, vT, _ = pullback_tangent_sharded(eng, v0, iters, iters, 0.0)
        out[0].copy_(u); out[1].copy_(s); out[2].copy_(vT)
        torch.cuda.synchronize()

    for i in range(min(Wm, 1)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(Wm, Wm + K):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax)
    h2d = 4 * S * (eng.n_in + (ctx.numel() if ctx is not None else 0) + k * eng.n_in)
    d2h = 4 * S * (k * eng.n_out + k + k * eng.n_in)

    # ---- per-kernel timing: two eagerly launched iterations with an event pair around every contraction launch ----
    prof, prof_iters = None, 2
    if rank == 0:
        pe = engS if S > 1 else eng                          # probe the configuration `value` was measured on
        if S > 1:
            for p in range(S):
                pe.set_point(xd[p], tval, ctxd, slot=p)
            pv0 = torch.cat(v0d[:S], 0)
        else:
            pe.set_point(xd[0], tval, ctxd)
            pv0 = v0d[0]
        torch.cuda.synchronize()
        pe.profile_begin()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        pe.pullback(pv0, prof_iters, prof_iters, 0.0)
        pe1.record()
        torch.cuda.synchronize()
        prof = pe.profile_read()
        prof["_eager_ms_per_iter"] = pe0.elapsed_time(pe1) / prof_iters

    secondary = None
    if wl == "sd15_mid_k5_i50" and not tangent and not args.no_secondary:
        del eng, engS
        torch.cuda.empty_cache()
        secondary = run_secondary(dev, rank, world)

    if rank == 0:
        pk_all = peaks()
        peak_tf, peak_burst, peak_hbm, peak_src = pk_all["sustained"], pk_all["burst"], pk_all["hbm"], pk_all["source"]
        flops_iter = 2.0 * k * f_tan * 1e9
        step_flops = flops_iter * iters + f_primal * 1e9
        achieved = step_flops * K * S / secs / 1e12 if f_tan else None
        kernels = {}
        for name in ("gemm_tc_kernel", "attn_lin_kernel"):
            kms, kfl, kn = prof[name]
            if kn:
                kernels[name] = {"launches_per_iter": kn / prof_iters, "ms_per_iter": kms / prof_iters,
                                 "avg_launch_us": 1e3 * kms / kn, "achieved_tflops": kfl / kms / 1e9,
                                 "share_of_eager_iter": kms / prof_iters / prof["_eager_ms_per_iter"],
                                 "frac": kfl / kms / 1e9 / pk_all["burst"]}
        if "attn_lin_kernel" in kernels:
            # what bounds it is the tcgen05 instruction rate at its small MMA shapes, not the tensor-pipe math rate: a JVP substep is
            # 6 MMAs of N = 64 and 8 of N = 48 = 971 clocks back to back for 384 clocks of peak-rate math (scripts/mma_rate.cu)
            kernels["attn_lin_kernel"]["instruction_roofline_frac_of_peak"] = 384.0 / 971.0
            kernels["attn_lin_kernel"]["instruction_roofline_source"] = "profiles/r3h_mma_rate.txt (one-CTA M = 128 MMAs, A from shared memory)"
        dom = max(kernels, key=lambda n: kernels[n]["ms_per_iter"]) if kernels else None
        # the GEMM launches by operand type, each against its own tensor-pipe peak (kind::tf32 = half the 16-bit rate)
        for name, pk in (("gemm_tc_kernel[kind::f16]", peak_burst), ("gemm_tc_kernel[kind::tf32]", pk_all["tf32_burst"])):
            kms, kfl, kn = prof.get(name, (0, 0, 0))
            if kn:
                kernels[name] = {"launches_per_iter": kn / prof_iters, "ms_per_iter": kms / prof_iters, "avg_launch_us": 1e3 * kms / kn,
                                 "achieved_tflops": kfl / kms / 1e9, "peak_tflops": pk, "frac": kfl / kms / 1e9 / pk}
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if dom and os.path.exists(tp):
            tj = json.load(open(tp)).get(wl, {}).get(dom)
            if tj:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "strong" if tangent else "weak", "vs_baseline": None, "dtype": "f16 tangents and operands, f32 accumulate (primal pass tf32)", "data": "synthetic",
                "config": {"workload": wl, "model": model_name + " (random-init)", "latent": [cfg["in_channels"], size, size], "op": op,
                           "block_idx": bi, "pca_rank": k, "power_iters": iters, "t": tval, "problems_per_rank": K * S, "slots": S,
                           "parallelism": (f"tangent-sharded x{world} (k columns of one problem split over the ranks, one all-gather of W per iteration)"
                                           if tangent else f"problem-sharded x{world}") +
                                          (f"; {S} problems per step batched through pb_set_slots (value, e2e and the kernel probes); `slots1` = one problem per call" if S > 1 else ""), "l2": "working set per iteration (>= 2.8 GB of weights + activations) exceeds the 126 MB L2",
                           "column_iters_per_s": value * k},
                "clocks": clocks,
                "e2e": {"value": (world if not tangent else 1) * K * S * iters / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches,
                "slots1": slots1,
                "roofline": {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["achieved_tflops"] if dom else None,
                             "peak": peak_burst, "unit": "TFLOP/s",
                             "frac": (kernels[dom]["achieved_tflops"] / peak_burst) if dom else None,
                             "peak_sustained": peak_tf, "peak_tf32": pk_all["tf32_burst"], "peak_tf32_source": pk_all["tf32_source"],
                             "traffic": traffic, "traffic_source": traffic_src,
                             "step_achieved": achieved, "step_frac": (achieved / peak_tf) if achieved else None,
                             "kernels": kernels,
                             "note": "achieved = algorithmic flops (2 M N K per product) of the dominant kernel's launches / their summed "
                                     "device time, event pair around every launch of 2 eager iterations; the kernel runs tcgen05 kind::f16 where the operand "
                                     "can be stored as halves (every 3x3 conv, GN/LN/GEGLU-fed linears) and kind::tf32 elsewhere (hardware peak = "
                                     "half of the bf16 peak used as denominator), fp32 accumulation; step_* = algorithmic "
                                     "flops of one whole step (50 x 2 k F_tan + primal, BASELINE.md s.3) / device time of the step; traffic = "
                                     "DRAM bytes per launch (ncu, profiles/); peak = " + peak_src}}
        if secondary is not None:
            line["secondary"] = secondary
        if world == 1 and not args.no_ref_gpu:
            # the reference algorithm on THIS GPU (torch eager autograd), one iteration at the workload's rank: the anchor of the
            # north star's ">= 20x the reference's single-GPU wall clock" (A100-equivalent taken as r = 1)
            rg = {}
            for tag, tf32 in (("fp32", False), ("tf32", True)):
                sec = gpu_reference_iteration(model_name, op, bi, k, tf32)
                rg[tag] = {"value": 1.0 / sec, "unit": UNIT, "s_per_iter": sec, "speedup_value": value / (1.0 / sec),
                           "speedup_e2e": line["e2e"]["value"] / (1.0 / sec)}
            rg["sample"] = (f"1 of {iters} subspace iterations at rank {k} on the full {model_name} {op}-{bi} problem, oracle port of "
                            "utils.py:722-816 through torch eager autograd on cuda:0 (cuDNN / cuBLAS), after one rank-1 warm-up iteration; "
                            "fp32 = TF32 off (the reference's precision), tf32 = allow_tf32 on")
            rg["device"] = torch.cuda.get_device_name(0)
            line["reference_gpu"] = rg
        if world == 1 and not args.no_cpu_baseline:
            k_s = k
            sec_iter = cpu_reference_iteration(model_name, op, bi, k_s)
            line["cpu_baseline"] = {"value": 1.0 / sec_iter, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"1 of {iters} subspace iterations at rank {k_s} on the full {model_name} {op}-{bi} problem "
                                              f"(oracle port of utils.py:722-816, torch-cpu fp32, {sec_iter:.1f} s)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sd15_mid_k5_i50", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` record (BASELINE configs[3]: rank 16, 10 timesteps, mixed schedule)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-this-GPU leg (torch eager autograd, 1 iteration)")
    ap.add_argument("--slots", type=int, default=0,
                    help="problem slots (pb_set_slots): each step solves this many independent problems as ONE tangent batch "
                         "(0 = default: min(5, 64 // pca_rank); 1 = one problem per step, the reference's granularity, which the "
                         "default run also reports as `slots1`)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu (the reference arm) or cuda (torch eager autograd on this box's GPU, SURVEY s.8d-ii)")
    ap.add_argument("--shard", default="problem", choices=["problem", "tangent"],
                    help="N > 1: independent problems per rank (default, weak scaling) or the k columns of each problem split over the ranks")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args, args.workload)
    else:
        if args.warmup < 3:
            args.warmup = 3                                   # timing rule: W >= 3
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()

"""Mint golden vectors for the pullback hot path by running the reference's OWN functions
(`/root/reference/src/utils/utils.py:722-816`, `:165-249`, `:438-527`, `:114-163`) verbatim,
bound with types.MethodType onto the restated diffusers-0.11.0 U-Net (oracle/unet_torch.py),
torch-CPU fp32.  Only runs in the authoring container (needs /root/reference); the resulting
small fixtures are committed under tests/golden/ and travel to the GPU box.

    python scripts/make_golden.py --case sd_tiny_mid [--case ...] | --all-small | --list
"""
import argparse
import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pullback_oracle as PO      # noqa: E402
from oracle import reference_shim as RS       # noqa: E402
from oracle import unet_torch as UT           # noqa: E402

#        name             config        op    blk k  iters
CASES = {
    "sd_tiny_mid":     ("sd_tiny",     "mid", 0, 3, 6),
    "sd_tiny_up1":     ("sd_tiny",     "up",  1, 3, 6),
    "sd_tiny_up0":     ("sd_tiny",     "up",  0, 2, 4),
    "sd_tiny_lin_mid": ("sd_tiny_lin", "mid", 0, 3, 6),
    "uncond_tiny_mid": ("uncond_tiny", "mid", 0, 2, 6),
    "sd_small_mid":    ("sd_small",    "mid", 0, 5, 8),
    "sd_small_up1":    ("sd_small",    "up",  1, 5, 4),
    # full-size BASELINE.json configs (minutes to hours of CPU)
    "sd15_mid_k5_i3":  ("sd15",        "mid", 0, 5, 3),
    "sd15_mid_k5_i50": ("sd15",        "mid", 0, 5, 50),
    "sd15_up1_k5_i2":  ("sd15",        "up",  1, 5, 2),
    "celebahq_mid_k2_i10": ("celebahq", "mid", 0, 2, 10),
    "sd21_768_mid_k5_i2": ("sd21_768", "mid", 0, 5, 2),
}
SMALL = [c for c in CASES if not c.startswith(("sd15", "celebahq", "sd21"))]


def run(case, out_dir):
    cfg_name, op, bi, k, iters = CASES[case]
    m = RS.bind(UT.build_unet(cfg_name, build_up=(op == "up")))
    n_params = sum(p.numel() for p in m.parameters())
    x, t, ctx = UT.synthetic_inputs(cfg_name)
    n_in = x[0].numel()
    # the reference draws V0 from the global RNG right after computing h_shape (utils.py:750)
    torch.manual_seed(0)
    v0 = PO.initial_subspace(n_in, k)
    torch.manual_seed(0)
    t0 = time.time()
    buf = io.StringIO()
    with torch.no_grad(), contextlib.redirect_stdout(buf):
        if ctx is not None:
            u, s, vT = m.local_encoder_pullback_zt(
                x, t, ctx, op=op, block_idx=bi, pca_rank=k, chunk_size=5,
                min_iter=iters, max_iter=iters, convergence_threshold=0.)
        else:
            u, s, vT = m.local_encoder_pullback_xt(
                x, t, op=op, block_idx=bi, pca_rank=k, chunk_size=25,
                min_iter=iters, max_iter=iters, convergence_threshold=0.)
    dt = time.time() - t0
    g = {"case": case, "config": cfg_name, "op": op, "block_idx": bi, "k": k, "iters": iters,
         "n_params": n_params, "v0": v0.clone(), "s": s.clone(), "vT": vT.clone().contiguous(),
         "n_out": u.shape[0], "u_norm": u.norm(dim=0).clone(),
         "ref_seconds": dt, "ref_threads": torch.get_num_threads(), "torch": str(torch.__version__),
         "ref_log": buf.getvalue()[-2000:]}
    u = u.contiguous()
    if u.numel() * 4 <= 4 << 20:
        g["u"] = u.clone()
    else:                       # large feature maps: keep every 16th row
        g["u_stride"] = 16
        g["u"] = u[::16].clone()
    os.makedirs(out_dir, exist_ok=True)
    torch.save(g, os.path.join(out_dir, case + ".pt"))
    print(f"{case}: n_params={n_params/1e6:.1f}M s={s.tolist()} {dt:.1f}s "
          f"({dt/iters:.2f} s/iter, {torch.get_num_threads()} threads)", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", action="append", default=[])
    ap.add_argument("--all-small", action="store_true")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    if a.list:
        print("\n".join(CASES)); sys.exit(0)
    for c in (SMALL if a.all_small else a.case):
        run(c, a.out)

"""Run-to-run agreement of the device path (eager vs eager, graph vs graph, eager vs graph)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_pullback_b200 as PB
from diffusion_pullback_b200 import synthetic as SY
name = sys.argv[1] if len(sys.argv) > 1 else "sd_small"
dev = "cuda:0"
unet = SY.SyntheticUNet(name, upto=("mid", 0), device=dev)
x, t, ctx = SY.synthetic_inputs(name)
size = unet.config["sample_size"]
eng = PB.PullbackEngine(PB.unet_config(unet), size, size, "mid", 0, 4, unet.config["ctx_len"], dev)
eng.bind(unet.state_dict())
torch.manual_seed(0)
q, _ = torch.linalg.qr(torch.randn(eng.n_in, 4)); V0 = q.T.contiguous().to(dev)
eng.set_point(x, float(t), ctx)
U1 = eng.jvp(V0); U2 = eng.jvp(V0)
print("jvp twice rel diff", float((U1 - U2).norm() / U1.norm()))
W1 = eng.vjp(U1); W2 = eng.vjp(U1)
print("vjp twice rel diff", float((W1 - W2).norm() / W1.norm()))
res = {}
for mode in ("eager", "eager", "graph", "graph"):
    eng.set_option("use_graph", int(mode == "graph"))
    eng.set_point(x, float(t), ctx)
    u, s, vT, _ = eng.pullback(V0, 6, 6, 0.0)
    print(mode, s.tolist())

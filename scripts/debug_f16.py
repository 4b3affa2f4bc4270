"""Debug: JVP / VJP of one config with the fp16-operand policy restricted by PB_F16_MASK, against the TF32 path."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import diffusion_pullback_b200 as PB
    from diffusion_pullback_b200 import synthetic as SY
    name, op, bi, k = sys.argv[2], sys.argv[3], int(sys.argv[4]), 3
    dev = torch.device("cuda:0")
    unet = SY.SyntheticUNet(name, upto=(op, bi), device=dev)
    size = unet.config["sample_size"]
    eng = PB.PullbackEngine(PB.unet_config(unet), size, size, op, bi, k, unet.config["ctx_len"], dev)
    eng.bind(unet.state_dict())
    x, t, ctx = SY.synthetic_inputs(name)
    eng.set_point(x, float(t), ctx)
    torch.manual_seed(0)
    V = torch.randn(k, eng.n_in, device=dev) / eng.n_in ** 0.5
    U = eng.jvp(V)
    G = torch.randn(k, eng.n_out, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    W = eng.vjp(G)
    torch.save((U.cpu(), W.cpu()), f"/tmp/f16dbg_{os.environ.get('PB_F16_MASK')}.pt")
else:
    import torch
    name, op, bi = sys.argv[1], sys.argv[2], sys.argv[3]
    outs = {}
    for mask in (0, 1, 2, 4, 8, 15):
        env = dict(os.environ, PB_F16_MASK=str(mask))
        subprocess.run([sys.executable, __file__, "child", name, op, bi], env=env, check=True)
        outs[mask] = torch.load(f"/tmp/f16dbg_{mask}.pt")
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    for mask in (1, 2, 4, 8, 15):
        print(name, op, bi, "mask", mask, "rel U", rel(outs[mask][0], outs[0][0]), "rel W", rel(outs[mask][1], outs[0][1]), flush=True)

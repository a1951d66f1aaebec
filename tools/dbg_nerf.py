import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semnerf_b200 import synth, build
from semnerf_b200.trainer import Trainer, default_cfgs
build.build()
DEV = "cuda"
cfgs = default_cfgs("nerf", n_samples=64, sc_lambda=0.0)
rays, extras = synth.make_rays(1024, seed=0)
rgbs, _, _ = synth.make_targets(rays, 0, seed=0)
batch = {"rays": rays.to(DEV), "extras": extras.to(DEV), "rgbs": rgbs.to(DEV)}
for kw in (dict(direct=True), dict(direct=False), dict(fused_loss=False)):
    tr = Trainer(cfgs, "nerf", 0, device=DEV, seed=0, **kw)
    p0 = tr.pbuf.clone()
    for i in range(3):
        l = tr.training_step(batch, epoch=3).item()
        g = tr.gbuf
        print(kw, i, "loss", l, "grad nan", int(torch.isnan(g).sum()), "inf", int(torch.isinf(g).sum()), "norm", float(g[~torch.isnan(g)].norm()),
              "param nan", int(torch.isnan(tr.pbuf).sum()), "m nan", int(torch.isnan(tr.exp_avg).sum()))
        if torch.isnan(g).any():
            m = tr.models["coarse"]
            for name, off, shape in m.table:
                n = 1
                for s in shape: n *= s
                c = int(torch.isnan(g[256 + off: 256 + off + n]).sum())
                if c: print("   nan in", name, c, "of", n)
            break

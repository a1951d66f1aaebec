#!/bin/bash
O=gpurun_out
T=${1:-r2r}
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${T}_smoke.log
cat > /tmp/san.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from semnerf_b200 import synth
from semnerf_b200.trainer import Trainer, default_cfgs
dev = "cuda"
for kind, C in (("semantic", 6), ("satnerf", 0), ("nerf", 0)):
    cfgs = default_cfgs(kind, n_samples=16, sc_lambda=0.0 if kind == "nerf" else 0.05, use_car_reg_loss=True, car_reg_loss_start=0,
                        use_beta_for_s=(kind == "semantic"), use_separate_beta_for_s=(kind == "semantic"), use_tj_for_s=(kind == "semantic"))
    tr = Trainer(cfgs, kind, C, device=dev, car_index=4, seed=0)
    n = 333
    rays, extras = synth.make_rays(n, seed=0)
    rgbs, labels, depths = synth.make_targets(rays, max(C, 1), seed=0)
    b = {"rays": rays.to(dev), "extras": extras.to(dev), "rgbs": rgbs.to(dev), "semantic": labels.to(torch.uint8).view(-1, 1).to(dev),
         "semantic_sparsity_mask": (torch.rand(n) < 0.7).to(dev)}
    d = {"rays": rays[:100].to(dev), "extras": extras[:100].to(dev), "depths": depths[:100].view(-1, 1).to(dev), "weights": torch.ones(100, 1, device=dev)}
    for i in range(2):
        l = tr.training_step(b, epoch=3, depth_batch=d if kind != "nerf" else None)
    img = tr.render_image(rays.to(dev), extras.to(dev), chunk=100)
    torch.cuda.synchronize()
    print(kind, "ok", float(l))
PY
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" $O/${T}_memcheck.log; tail -8 $O/${T}_memcheck.log

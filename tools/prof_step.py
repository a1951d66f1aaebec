"""Run a few training steps of the bench workload (for ncu captures of the kernels inside a real step).

    python tools/prof_step.py [rays] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semnerf_b200 import build, synth
from semnerf_b200.trainer import Trainer, default_cfgs

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
cfgs = default_cfgs("semantic", n_samples=64, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=0)
tr = Trainer(cfgs, "semantic", 6, device=dev, car_index=4, seed=0)
rays, extras = synth.make_rays(B, seed=0)
rgbs, labels, _ = synth.make_targets(rays, 6, seed=0)
batch = {k: v.to(dev) for k, v in {"rays": rays, "extras": extras, "rgbs": rgbs, "semantic": labels}.items()}
for i in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = tr.training_step(batch, epoch=3)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {i}: {e0.elapsed_time(e1):.2f} ms loss {loss.item():.4f}")

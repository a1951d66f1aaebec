"""Host time of the eager (no CUDA graph) direct step against its device time: is the launch sequence host-bound anywhere?

    python tools/host_time.py [rays] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semnerf_b200 import build, synth
from semnerf_b200.trainer import Trainer, default_cfgs

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
cfgs = default_cfgs("semantic", n_samples=64, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=0)
for graph in (False, True):
    tr = Trainer(cfgs, "semantic", 6, device=dev, car_index=4, seed=0, graph=graph)
    rays, extras = synth.make_rays(B, seed=0)
    rgbs, labels, _ = synth.make_targets(rays, 6, seed=0)
    batch = {k: v.to(dev) for k, v in {"rays": rays, "extras": extras, "rgbs": rgbs, "semantic": labels}.items()}
    for i in range(5):
        tr.training_step(batch, epoch=3)
    torch.cuda.synchronize()
    host = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all = time.perf_counter()
    e0.record()
    for i in range(steps):
        t0 = time.perf_counter()
        tr.training_step(batch, epoch=3)
        host.append(time.perf_counter() - t0)
    e1.record()
    t_issue = time.perf_counter() - t_all
    torch.cuda.synchronize()
    host.sort()
    print(f"HOST_TIME graph={graph} rays={B}: device {e0.elapsed_time(e1) / steps:.3f} ms/step; host call median "
          f"{1e3 * host[len(host) // 2]:.3f} ms, min {1e3 * host[0]:.3f}, max {1e3 * host[-1]:.3f}; all {steps} calls issued in "
          f"{1e3 * t_issue:.1f} ms")
    del tr

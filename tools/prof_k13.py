"""Time K1 (sample + encode) and K3 (composite forward / backward) alone at a training-size and a
render-chunk-size ray count and report achieved ALGORITHMIC GB/s against the measured HBM peak.

    python tools/prof_k13.py [n_rays ...]

Algorithmic bytes:  K1: 48 B/ray read + per sample z 4 + enc 2*enc_ld (x2 with the solar row)
+ aux 32 written.  K3 forward: S*(4*(9+C) + 4 + 8) + 12 + 4 + 4C + 8 per ray.  K3 backward: per sample
read 4*(9+C) + 4 + g_weights 4 + g_direct 4*(9+C), write 4*(9+C); + per-ray grads.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semnerf_b200 import _lib, build, synth
from semnerf_b200._lib import check, ptr, stream
from semnerf_b200.autograd import encode_rays
from semnerf_b200.model import RSSemanticNeRFB200

build.build()
lib = _lib.load()
dev = torch.device("cuda", 0)
S, C = 64, 6
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


sizes = [int(a) for a in sys.argv[1:]] or [8192, 40960]
from semnerf_b200.trainer import default_cfgs
cfgs = default_cfgs("semantic", n_samples=S, sc_lambda=0.05)
import types
model = RSSemanticNeRFB200(cfgs, types.SimpleNamespace(semantic_n_classes=C)).to(dev)
emb = torch.nn.Embedding(50, 4).to(dev)
n_out = 9 + C
for n in sizes:
    rays, extras = synth.make_rays(n, seed=1)
    rays, extras = rays.to(dev), extras.to(dev)
    P = n * S
    # K1 through the C ABI directly with preallocated outputs (the Python wrapper's allocations and
    # small tensor ops cost more host time than the kernel runs); 4 back-to-back launches per timing
    from semnerf_b200.autograd import t_steps
    zv = torch.empty(n, S, device=dev); enc = torch.empty(P, model.enc_ld, dtype=torch.bfloat16, device=dev)
    enc_sc = torch.empty_like(enc); aux = torch.empty(P, 16, dtype=torch.bfloat16, device=dev)
    ts = t_steps(S, dev); ew = emb.weight.detach().contiguous()
    k1_args = (ptr(rays), ptr(extras), None, 3, None, 0, ptr(ts), ptr(ew), 50, 4, None, None, None, None, 0, n, S, model.kind, 0,
               ptr(zv), ptr(enc), ptr(enc_sc), ptr(aux), None, stream())

    def k1x4():
        for _ in range(4):
            check(lib.snb_sample_encode(*k1_args), "k1")
    t = timed(k1x4) / 4
    k1_bytes = n * 48 + P * (4 + 2 * model.enc_ld * 2 + 32)
    print(f"K1 sample+encode (main + solar rows) n={n}: {t * 1e6:8.1f} us  {k1_bytes / t / 1e9:8.1f} GB/s  "
          f"= {k1_bytes / t / 1e9 / HBM:.3f} of measured HBM peak ({k1_bytes / 1e6:.1f} MB)")
    out = torch.rand(P, n_out, device=dev)
    z = torch.sort(torch.rand(n, S, device=dev), dim=1).values
    rgb = torch.empty(n, 3, device=dev); depth = torch.empty(n, device=dev)
    w = torch.empty(n, S, device=dev); T = torch.empty(n, S, device=dev)
    sem = torch.empty(n, C, device=dev); lab = torch.empty(n, dtype=torch.int64, device=dev)
    t = timed(lambda: check(lib.snb_composite_forward(ptr(out), ptr(z), n, S, n_out, C, 0, ptr(rgb), ptr(depth), ptr(w), ptr(T),
                                                      ptr(sem), ptr(lab), stream()), "k3f"))
    k3f = n * (S * (4 * n_out + 4 + 8) + 12 + 4 + 4 * C + 8)
    print(f"K3 composite forward n={n}:              {t * 1e6:8.1f} us  {k3f / t / 1e9:8.1f} GB/s  = {k3f / t / 1e9 / HBM:.3f} "
          f"of measured HBM peak ({k3f / 1e6:.1f} MB)")
    g_rgb = torch.rand(n, 3, device=dev); g_d = torch.rand(n, device=dev); g_w = torch.rand(n, S, device=dev)
    g_sem = torch.rand(n, C, device=dev); g_dir = torch.rand(P, n_out, device=dev); g_out = torch.empty(P, n_out, device=dev)
    t = timed(lambda: check(lib.snb_composite_backward(ptr(out), ptr(z), n, S, n_out, C, 0, ptr(g_rgb), ptr(g_d), ptr(g_w), None,
                                                       ptr(g_sem), ptr(g_dir), ptr(g_out), stream()), "k3b"))
    k3b = n * (S * (4 * n_out + 4 + 4 + 4 * n_out + 4 * n_out) + 12 + 4 + 4 * C)
    print(f"K3 composite backward n={n}:             {t * 1e6:8.1f} us  {k3b / t / 1e9:8.1f} GB/s  = {k3b / t / 1e9 / HBM:.3f} "
          f"of measured HBM peak ({k3b / 1e6:.1f} MB)")

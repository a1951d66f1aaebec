"""K1 / K3 alone against the measured HBM copy peak (the bench's own harness), at the training batch, the render chunk and
4 chunks.    python tools/prof_k13.py [rays ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from semnerf_b200 import _lib, build
build.build()
lib = _lib.load()
dev = torch.device("cuda", 0)
peak_tf, peak_hbm, which = bench.peaks()
for n in [int(a) for a in sys.argv[1:]] or [8192, 40960, 163840]:
    r = bench.hbm_kernel_rooflines(lib, dev, peak_hbm, n)
    print(n, json.dumps({k: {"frac": round(v["frac"], 3), "us": round(v["us"], 1), "GBps": round(v["achieved"])} for k, v in r.items() if isinstance(v, dict)}))

"""BASELINE.json configs[2..4] on 1..8 GPUs (bench.py measures configs[1], the headline): one JSON line per config.

    python tools/bench_configs.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_configs.py [--views 2] [--skip-train]

  config 3  full-image render: 800 x 800 synthetic view, S = 64, semantic model, rgb / depth / label; the image's rays are
            sharded contiguously across the ranks, no communication; samples/s = 800*800*64 / max-over-ranks time
  config 4  data-parallel training at a 65 536-ray GLOBAL batch (strong scaling over the ranks) with the bucketed NCCL
            gradient all-reduce; rays/s
  config 5  dense eval sweep: depth + rgb point-cloud extraction at S = 128, `--views` synthetic 798 x 758 views per
            rank (views sharded across ranks, no communication); samples/s
  fp32      the fp32 verification mode on one 40 960-ray chunk (rank 0 only); samples/s
All device-timed (CUDA events), max over ranks; inputs resident in HBM; synthetic rays (DFC2019 is not available offline).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from semnerf_b200 import build, dist as snb_dist, synth
from semnerf_b200.pointcloud import extract_pointcloud
from semnerf_b200.trainer import Trainer, default_cfgs


def timed_max(fn, dev, world):
    snb_dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    return ms.item() * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=2, help="config 5: 798x758 views per rank")
    ap.add_argument("--skip-train", action="store_true")
    ap.add_argument("--train-steps", type=int, default=5)
    args = ap.parse_args()
    rank, local, world = snb_dist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        build.build()
    snb_dist.barrier()
    C = 6
    lines = []

    # ---- config 3: 800 x 800 render, rays sharded across ranks ------------------------------------------------
    cfgs = default_cfgs("semantic", n_samples=64, sc_lambda=0.05)
    tr = Trainer(cfgs, "semantic", C, device=dev, world=1, rank=0, seed=0)
    n_img = 800 * 800
    lo, hi = snb_dist.shard_range(n_img, rank, world)
    rays, extras = synth.make_rays(hi - lo, seed=1000 + rank)
    rays, extras = rays.to(dev), extras.to(dev)
    tr.render_image(rays[:40960], extras[:40960])   # warm-up (allocations, first launches)
    res = {}

    def render3():
        res["img"] = tr.render_image(rays, extras)
    t = timed_max(render3, dev, world)
    lines.append({"config": "3: 800x800 render (rgb, depth, semantic label), S=64, rays sharded, no communication",
                  "metric": "render_samples_per_s", "value": n_img * 64 / t, "unit": "samples/s", "n_gpus": world,
                  "seconds": t, "passes": "main pass (render_image skips the solar-correction pass: nothing it returns needs it)"})

    def render3_sc():
        res["img"] = tr.render_image(rays, extras, keys=("rgb_coarse", "depth_coarse", "semantic_label_coarse", "sun_sc_coarse"))
    t = timed_max(render3_sc, dev, world)
    lines.append({"config": "3: 800x800 render, main + solar-correction pass (what the reference's batched_inference computes)",
                  "metric": "render_samples_per_s", "value": n_img * 64 / t, "unit": "samples/s", "n_gpus": world, "seconds": t})

    # ---- config 5: dense eval sweep, S = 128, depth + rgb, views sharded across ranks ---------------------------
    cfg5 = default_cfgs("semantic", n_samples=128, sc_lambda=0.0)
    tr5 = Trainer(cfg5, "semantic", C, device=dev, world=1, rank=0, seed=0)
    n_view = 798 * 758
    vr, ve = synth.make_rays(n_view, seed=2000 + rank)
    vr, ve = vr.to(dev), ve.to(dev)
    extract_pointcloud(tr5.renderer, tr5.models, vr[:40960], ve[:40960], center=(0.0, 0.0, 0.0), scale=100.0)

    def sweep():
        for _ in range(args.views):
            res["pc"] = extract_pointcloud(tr5.renderer, tr5.models, vr, ve, center=(0.0, 0.0, 0.0), scale=100.0)
    t = timed_max(sweep, dev, world)
    lines.append({"config": "5: depth + rgb point-cloud extraction, S=128, 798x758 views sharded across ranks",
                  "metric": "render_samples_per_s", "value": world * args.views * n_view * 128 / t, "unit": "samples/s",
                  "n_gpus": world, "views": world * args.views, "seconds": t,
                  "full_sweep_estimate_s_71_views": 71 * n_view * 128 / (world * args.views * n_view * 128 / t)})

    # ---- fp32 verification mode, one chunk (rank 0) ---------------------------------------------------------------
    if rank == 0:
        tr.renderer.cfgs, tr.cfgs = cfgs, cfgs
        n32 = 40960
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            tr.renderer.render_rays(tr.models, rays[:4096], extras[:4096], render_options={"precision": "fp32"})
            torch.cuda.synchronize()
            e0.record()
            tr.renderer.render_rays(tr.models, rays[:n32], extras[:n32], render_options={"precision": "fp32", "seed": 1})
            e1.record()
            torch.cuda.synchronize()
        t32 = e0.elapsed_time(e1) * 1e-3
        lines.append({"config": "fp32 verification mode, 40960-ray chunk, S=64, main + solar pass (CUDA cores, fp32 FMA)",
                      "metric": "render_samples_per_s", "value": n32 * 64 / t32, "unit": "samples/s", "n_gpus": 1,
                      "seconds": t32})
    snb_dist.barrier()
    del tr5, vr, ve, res
    torch.cuda.empty_cache()

    # ---- config 4: DP training, 65 536-ray global batch ----------------------------------------------------------
    if not args.skip_train:
        G = 65536
        B = G // world
        # ~25 KB of saved activations per sample and pass: a rank's share runs as 8192-ray micro-batches that accumulate into
        # one gradient, ONE all-reduce + optimiser step per global batch (one GPU: 8 micro-batches, 8 GPUs: 1)
        cfg4 = default_cfgs("semantic", n_samples=64, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=0)
        tr4 = Trainer(cfg4, "semantic", C, device=dev, car_index=4, world=world, rank=rank, seed=0, micro_batch=8192)
        br, be = synth.make_rays(B, seed=3000 + rank)
        rgbs, labels, _ = synth.make_targets(br, C, seed=rank)
        batch = {k: v.to(dev) for k, v in {"rays": br, "extras": be, "rgbs": rgbs, "semantic": labels}.items()}
        for _ in range(2):
            tr4.training_step(batch, epoch=3, ray_offset=rank * B, global_rays=G)

        def train():
            for _ in range(args.train_steps):
                tr4.training_step(batch, epoch=3, ray_offset=rank * B, global_rays=G)
        t = timed_max(train, dev, world)
        lines.append({"config": "4: data-parallel semantic-NeRF training, 65536-ray GLOBAL batch, NCCL gradient all-reduce",
                      "metric": "train_rays_per_s", "value": G * args.train_steps / t, "unit": "rays/s", "n_gpus": world,
                      "rays_per_gpu_per_step": B, "steps": args.train_steps, "ms_per_step": t / args.train_steps * 1e3,
                      "scaling": "strong"})
    if rank == 0:
        for ln in lines:
            print(json.dumps(ln), flush=True)
    snb_dist.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

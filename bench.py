#!/usr/bin/env python
"""Benchmark of the ray-rendering hot path (BASELINE.json: training rays/s, render samples/s,
MLP tensor-pipe fraction of peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch RAYS_PER_GPU] [--impl reference]

A "step" is one optimisation step of the Semantic-NeRF pipeline (BASELINE.json configs[1]: semantic
model, 6 classes, transient/car regularisation, solar-correction pass, bf16) on a synthetic batch of
rays: sample + encode -> MLP (main + solar pass) -> composite -> losses -> backward -> Adam.
For N > 1 launch under torchrun (one rank per GPU, NCCL); per-GPU work is fixed ("weak" scaling).
Prints ONE JSON line on rank 0.

--impl reference times the CPU port of the reference's path (oracle/) on the host cores; the
reference itself is pure Python that cannot travel to the GPU box (no /root/reference there).
"""
from __future__ import annotations

import argparse
import ctypes as C
import ctypes as C_
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLASSES, N_SAMPLES, CAR_INDEX = 6, 64, 4
# algorithmic MLP FLOPs per training ray (semantic, C=6, S=64): main pass 16.862 MFLOP/sample
# (fwd + dgrad + wgrad, no dgrad into the inputs) + solar pass 14.470 MFLOP/sample.
ALG_FLOP_PER_TRAIN_RAY = 64 * (16_862_208 + 14_470_144)
ALG_FLOP_PER_RENDER_SAMPLE = 5_641_216 + 4_844_544  # all heads + the solar pass forward
# dram__bytes_read.sum + dram__bytes_write.sum of the GEMM kernels of one 8192-ray step, from the ncu --set full capture
# under profiles/ (total printed by tools/ncu_summary.py on the capture tools/round_profile.sh takes)
NCU_TRAFFIC_BYTES_PER_STEP = 67.932e9   # profiles/r01g_ncu_step_gemms.csv (34 launches of one step)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi polled every 50 ms from BEFORE the warm-up (its start-up takes longer than a short timed region); the
    samples reported are those whose timestamps fall inside the timed region marked with begin() / end()."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = f"/tmp/snb_clocks_{os.getpid()}.csv"
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)                     # let the sample that covers the end of the region land
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        import datetime
        rows = []
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[2]), float(parts[3]), float(parts[4]), parts[5:9]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.03 <= r[0] <= (self.t1 or r[0]) + 0.06]
        use = inside if inside else rows     # a region shorter than the polling period: fall back to every sample taken
        sm = sorted(r[1] for r in use)
        reasons = set()
        for r in use:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        pw = sorted(r[3] for r in use)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": use[0][2] if use else None, "reasons": sorted(reasons),
                "samples": len(use), "samples_in_timed_region": len(inside), "power_w_median": pw[len(pw) // 2] if pw else None}


def make_batch(n, seed, pinned=False):
    from semnerf_b200 import synth
    rays, extras = synth.make_rays(n, seed=seed)
    rgbs, labels, _ = synth.make_targets(rays, N_CLASSES, seed=seed)
    b = {"rays": rays, "extras": extras, "rgbs": rgbs, "semantic": labels}
    if pinned:
        b = {k: v.pin_memory() for k, v in b.items()}
    return b


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, one optimisation step on a bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_training_steps(n_rays: int, steps: int, warmup: int):
    from oracle import render_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    spec = O.ModelSpec(kind="semantic", n_classes=N_CLASSES)
    params, emb = O.make_params(spec, seed=0)
    params = {k: v.requires_grad_(True) for k, v in params.items()}
    emb = emb.requires_grad_(True)
    opt = torch.optim.Adam(list(params.values()) + [emb], lr=5e-4)
    b = make_batch(n_rays, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        u = torch.rand(n_rays, N_SAMPLES)
        res = O.render_rays(params, emb, spec, b["rays"], b["extras"], N_SAMPLES, u=u, sc_lambda=0.05)
        loss = O.satnerf_loss(res, b["rgbs"]) + O.semantic_loss(res, b["semantic"], ignore_index=CAR_INDEX) + \
            O.car_reg_loss(res, b["semantic"], CAR_INDEX)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return n_rays * len(times) / sum(times), torch.get_num_threads(), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_rays
    rps, cores, sec = cpu_training_steps(n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "train_rays_per_s", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{n} rays x {N_SAMPLES} samples per step (same step, bounded batch)"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, cpu=False, graph=False):
    c = {
        "workload": "Semantic-NeRF training step (BASELINE configs[1]): semantic 8x512 SIREN MLP, C=6, S=64, "
                    "solar-correction pass (sc_lambda=0.05), SatNerfLoss + SemanticLoss(ignore car) + "
                    "SemanticCarRegLoss, Adam; steady-state step (after the depth-supervision drop)",
        "rays_per_gpu_per_step": args.cpu_rays if cpu else args.batch, "samples_per_ray": N_SAMPLES,
        "n_classes": N_CLASSES, "parallelism": f"dp{args.gpus}",
    }
    if cpu:
        c["losses"] = "plain PyTorch on the host: the oracle's SatNerfLoss + SemanticLoss + SemanticCarRegLoss (autograd, torch Adam)"
        c["l2"] = "n/a (host cores)"
        c["note"] = ("bounded sample: the CPU arm steps fewer rays per step than the GPU arm (the metric is rays/s; "
                     "CPU rays/s is flat in the batch size)")
    else:
        c["losses"] = ("loss modules on render_rays()" if getattr(args, "module_losses", False)
                       else "fused into the compositing kernel")
        c["l2"] = "per-step activation working set (~30 GB at 8192 rays) >> 126 MB L2; no flush needed"
        c["step"] = ("direct kernel sequence on persistent buffers" + (", replayed as one CUDA graph" if graph else "")
                     if not getattr(args, "module_losses", False) else "autograd")
    return c


# ------------------------------------------------------------------------------------------------------
# HBM-bound kernels of the path (K1 sample + encode, K3 composite) timed alone against the measured copy peak
# ------------------------------------------------------------------------------------------------------
def hbm_kernel_rooflines(lib, dev, peak_hbm, n=40960):
    """Algorithmic bytes (DESIGN.md 4 / SURVEY 8d) / CUDA-event time per launch.  Every kernel runs over R independent
    replicas of its buffers back to back inside one event pair (R x footprint >= 512 MB >> 126 MB L2, so no launch finds its
    inputs cached and the launch queue never runs dry - the state the kernels run in inside a step); median of 3 rounds."""
    import types
    from semnerf_b200 import synth
    from semnerf_b200._lib import LossParams, check, ptr, stream
    from semnerf_b200.autograd import t_steps
    from semnerf_b200.model import RSSemanticNeRFB200
    from semnerf_b200.trainer import default_cfgs
    S, C = N_SAMPLES, N_CLASSES
    n_out, P = 9 + C, n * N_SAMPLES

    def timed(fns):
        for f in fns[:1]:
            f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for f in fns:
                f()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / len(fns))
        ts.sort()
        return ts[1] * 1e-3

    def reps(nbytes):
        return int(min(16, max(2, -(-512 * 1024 * 1024 // nbytes))))

    def entry(b, t, r, **kw):
        return {"bound": "hbm", "achieved": b / t / 1e9, "peak": peak_hbm, "unit": "GB/s", "frac": b / t / 1e9 / peak_hbm,
                "bytes": b, "us": t * 1e6, "replicas": r, **kw}

    cfgs = default_cfgs("semantic", n_samples=S, sc_lambda=0.05)
    model = RSSemanticNeRFB200(cfgs, types.SimpleNamespace(semantic_n_classes=C)).to(dev)
    emb = torch.nn.Embedding(50, 4).to(dev)
    rays, extras = synth.make_rays(n, seed=1)
    rays, extras = rays.to(dev), extras.to(dev)
    ts, ew = t_steps(S, dev), emb.weight.detach().contiguous()
    res = {}
    # ---- K1: sample + encode, main and solar rows ----
    b = n * 48 + P * (4 + 2 * model.enc_ld * 2 + 32)
    r = reps(b)
    bufs = [(torch.empty(n, S, device=dev), torch.empty(P, model.enc_ld, dtype=torch.bfloat16, device=dev),
             torch.empty(P, model.enc_ld, dtype=torch.bfloat16, device=dev), torch.empty(P, 16, dtype=torch.bfloat16, device=dev))
            for _ in range(r)]
    fns = [(lambda zv=zv, e=e, es=es, a=a: check(lib.snb_sample_encode(
        ptr(rays), ptr(extras), None, 3, None, 0, ptr(ts), ptr(ew), 50, 4, None, None, None, None, 0, n, S, model.kind, 0,
        ptr(zv), ptr(e), ptr(es), ptr(a), None, stream()), "k1")) for zv, e, es, a in bufs]
    res["k1_sample_encode"] = entry(b, timed(fns), r, note="write-only kernel (main + solar rows)")
    del bufs, fns
    # ---- K3 forward / backward / fused loss ----
    b_f = n * (S * (4 * n_out + 4 + 8) + 12 + 4 + 4 * C + 8)
    b_b = n * (S * (4 * n_out + 4 + 4 + 4 * n_out + 4 * n_out) + 12 + 4 + 4 * C)
    b_l = n * (S * (4 * n_out + 4 + 4 * n_out) + 12 + 8)
    r = reps(b_f)
    outs = [torch.rand(P, n_out, device=dev) for _ in range(r)]
    zs = [torch.sort(torch.rand(n, S, device=dev), dim=1).values for _ in range(r)]
    rgb, depth = torch.empty(n, 3, device=dev), torch.empty(n, device=dev)
    ws = [(torch.empty(n, S, device=dev), torch.empty(n, S, device=dev)) for _ in range(r)]
    sem, lab = torch.empty(n, C, device=dev), torch.empty(n, dtype=torch.int64, device=dev)
    fns = [(lambda o=o, z=z, w=w: check(lib.snb_composite_forward(ptr(o), ptr(z), n, S, n_out, C, 0, ptr(rgb), ptr(depth), ptr(w[0]),
                                                                 ptr(w[1]), ptr(sem), ptr(lab), stream()), "k3f"))
           for o, z, w in zip(outs, zs, ws)]
    res["k3_composite_forward"] = entry(b_f, timed(fns), r)
    g_rgb, g_d = torch.rand(n, 3, device=dev), torch.rand(n, device=dev)
    g_sem = torch.rand(n, C, device=dev)
    g_w = ws[0][0].uniform_()
    g_dirs = [torch.rand(P, n_out, device=dev) for _ in range(r)]
    g_outs = [torch.empty(P, n_out, device=dev) for _ in range(r)]
    fns = [(lambda o=o, z=z, gd=gd, go=go: check(lib.snb_composite_backward(ptr(o), ptr(z), n, S, n_out, C, 0, ptr(g_rgb), ptr(g_d),
                                                                          ptr(g_w), None, ptr(g_sem), ptr(gd), ptr(go), stream()), "k3b"))
           for o, z, gd, go in zip(outs, zs, g_dirs, g_outs)]
    res["k3_composite_backward"] = entry(b_b, timed(fns), r)
    del g_dirs
    # the training step's form: composite + losses + composite backward in one pass (snb_composite_loss, main pass)
    gt = torch.rand(n, 3, device=dev)
    labels = torch.randint(0, C, (n,), device=dev)
    counts = torch.tensor([float(n), float(n) / C, 0.0, 0.0], device=dev)
    terms = torch.zeros(8, device=dev)
    lp = LossParams(mode=0, color=1, beta_min=0.05, inv_n=1.0 / n, lambda_s=0.04, ignore_index=CAR_INDEX, lambda_c=0.1,
                    car_label=CAR_INDEX, lambda_sc=0.05, lambda_ds=0.0, flags=0)
    fns = [(lambda o=o, z=z, go=go: check(lib.snb_composite_loss(ptr(o), ptr(z), n, S, n_out, C, ptr(gt), ptr(labels), None, None,
                                                               None, ptr(counts), C_.byref(lp), ptr(go), ptr(terms), stream()), "k3l"))
           for o, z, go in zip(outs, zs, g_outs)]
    res["k3_composite_loss_fused"] = entry(b_l, timed(fns), r)
    res["rays"] = n
    return res


# ------------------------------------------------------------------------------------------------------
# the same step as eager cuBLAS + ATen on the SAME GPU (SURVEY 2.3 / 8d "kernel to beat"): the oracle's restatement of
# the reference path run on CUDA tensors - what the reference's own modules execute on a GPU, launch for launch
# ------------------------------------------------------------------------------------------------------
def gpu_eager_steps(dev, n_rays: int, mode: str, steps: int = 4, warmup: int = 2):
    """mode "tf32": torch.set_float32_matmul_precision("high") as in the reference's run template (run/run_template.toml:15);
    mode "bf16": the same under torch.autocast(bfloat16).  Returns rays/s (CUDA events)."""
    from oracle import render_oracle as O
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    try:
        spec = O.ModelSpec(kind="semantic", n_classes=N_CLASSES)
        params, emb = O.make_params(spec, seed=0)
        params = {k: v.to(dev).requires_grad_(True) for k, v in params.items()}
        emb = emb.to(dev).requires_grad_(True)
        opt = torch.optim.Adam(list(params.values()) + [emb], lr=5e-4)
        b = {k: v.to(dev) for k, v in make_batch(n_rays, seed=0).items()}

        def step():
            u = torch.rand(n_rays, N_SAMPLES, device=dev)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                res = O.render_rays(params, emb, spec, b["rays"], b["extras"], N_SAMPLES, u=u, sc_lambda=0.05)
            res = {k: (v.float() if v.is_floating_point() else v) for k, v in res.items()}
            loss = O.satnerf_loss(res, b["rgbs"]) + O.semantic_loss(res, b["semantic"], ignore_index=CAR_INDEX) + \
                O.car_reg_loss(res, b["semantic"], CAR_INDEX)
            opt.zero_grad()
            loss.backward()
            opt.step()
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return n_rays * steps / (e0.elapsed_time(e1) * 1e-3)
    finally:
        torch.set_float32_matmul_precision(prev)
        del params, emb
        torch.cuda.empty_cache()


def cpu_config1(n_rays: int = 1024):
    """BASELINE configs[0] exactly (SURVEY 8d "Config 1"): SatNeRF render_rays + (rgb.sum() + depth.sum()).backward(), 1024 rays,
    64 samples, no semantic head, on the host cores; sc_lambda 0 and 0.05; 1 warm-up + best of 3."""
    from oracle import render_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    spec = O.ModelSpec(kind="satnerf", n_classes=0)
    params, emb = O.make_params(spec, seed=0)
    params = {k: v.requires_grad_(True) for k, v in params.items()}
    rays, extras = O.synthetic_rays(n_rays, seed=0)
    out = {}
    for sc in (0.0, 0.05):
        best = None
        for i in range(4):
            t0 = time.perf_counter()
            u = torch.rand(n_rays, N_SAMPLES)
            res = O.render_rays(params, emb, spec, rays, extras, N_SAMPLES, u=u, sc_lambda=sc)
            loss = res["rgb_coarse"].sum() + res["depth_coarse"].sum()
            if sc > 0:   # the solar pass enters the graph through its outputs
                loss = loss + res["sun_sc_coarse"].sum()
            for v in params.values():
                v.grad = None
            loss.backward()
            dt = time.perf_counter() - t0
            if i > 0:
                best = dt if best is None else min(best, dt)
        out[f"sc_lambda_{sc}"] = {"rays_per_s": n_rays / best, "s_per_step": best}
    return {"workload": f"SatNeRF render_rays fwd+bwd, {n_rays} rays x {N_SAMPLES} samples, 8x512 SIREN, no semantic head, fp32",
            "cores": torch.get_num_threads(), "kind": "port", "timing": "1 warm-up + best of 3", **out}


def newest_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one step's GEMM kernels from the newest per-launch ncu capture under
    profiles/ (tools/ncu_summary.py writes the '# total' row)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_step_gemms.csv")))
    for f in reversed(files):
        for line in open(f):
            if line.startswith("# total"):
                parts = line.strip().split(",")
                try:
                    return float(parts[5]), os.path.relpath(f, ROOT)
                except (IndexError, ValueError):
                    break
    return None, None


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def timed_steps(fn, steps, dev, world):
    """device time of `steps` calls of fn(i), barrier + synchronize on both sides, max over ranks (ms)"""
    from semnerf_b200 import dist as snb_dist
    snb_dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = fn(i)
    e1.record()
    torch.cuda.synchronize()
    snb_dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    return ms.item(), out


def run_gpu(args):
    from semnerf_b200 import _lib, build, dist as snb_dist
    from semnerf_b200.trainer import Trainer, default_cfgs
    rank, local, world = snb_dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        build.build()
    snb_dist.barrier()
    lib = _lib.load()
    sampler = ClockSampler(local) if rank == 0 else None      # started before the warm-up: see ClockSampler
    B = args.batch
    graph = world == 1 and not args.no_graph and not args.module_losses
    cfgs = default_cfgs("semantic", n_samples=N_SAMPLES, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=0)
    tr = Trainer(cfgs, "semantic", N_CLASSES, device=dev, car_index=CAR_INDEX, world=world, rank=rank, seed=0,
                 fused_loss=not args.module_losses, graph=graph, micro_batch=8192)
    host = [make_batch(B, seed=100 * rank + i, pinned=True) for i in range(4)]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def step_resident(i):
        return tr.training_step(resident[i % len(resident)], epoch=3, ray_offset=rank * B)

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    torch.cuda.synchronize()
    # ---- timed region 1: inputs resident in HBM, device-timed ----------------------------------------
    lib.snb_profile_begin(0)
    if sampler:
        sampler.begin()
    ms_total, loss = timed_steps(step_resident, args.steps, dev, world)
    if sampler:
        sampler.end()
    clocks = sampler.stop() if sampler else None
    gl, tl = C.c_int64(), C.c_int64()
    lib.snb_profile_end(None, C.byref(gl), C.byref(tl), None)
    launches = tl.value
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- timed region 2: end to end through the public API, host buffers ------------------------------
    # every step copies its inputs from pinned host memory (inside training_step: the staging copies of the direct step)
    # and reads the loss back
    snb_dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = tr.training_step(host[i % len(host)], epoch=3, ray_offset=rank * B)
        loss_host = loss.item()          # device -> host read of the step's result
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t_e2e, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * args.steps / t_e2e.item()
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    # ---- strong scaling (BASELINE configs[3]): a fixed 65 536-ray GLOBAL batch, 65 536 / N rays per GPU in 8192-ray
    # micro-batches that accumulate into one gradient, ONE all-reduce + optimiser step per global batch ------------
    strong = None
    if not args.no_extras:
        G = 65536
        Bs = G // world
        sb = {k: v.to(dev) for k, v in make_batch(Bs, seed=900 + rank).items()}
        tr.training_step(sb, epoch=3, ray_offset=rank * Bs, global_rays=G)
        nst = 3
        ms_s, _ = timed_steps(lambda i: tr.training_step(sb, epoch=3, ray_offset=rank * Bs, global_rays=G), nst, dev, world)
        strong = {"scaling": "strong", "global_rays_per_step": G, "rays_per_gpu_per_step": Bs, "micro_batch_rays": 8192,
                  "value": G * nst / (ms_s * 1e-3), "unit": "rays/s", "ms_per_step": ms_s / nst, "steps": nst}
        del sb

    # ---- roofline of the dominant kernels (the tcgen05 GEMMs), timed live with CUDA events per launch ------
    # every rank runs these steps (they contain the gradient all-reduce); rank 0 reports its own launches
    roof = cpu = render = hbm = small = early = eager = cfg1 = None
    nprof = 4
    snb_dist.barrier()
    lib.snb_profile_begin(1)
    use_graph, tr.use_graph = tr.use_graph, False     # per-launch event timing needs the launches to go through the host
    for i in range(nprof):
        step_resident(i)
    tr.use_graph = use_graph
    gms, gl2, tl2, macs = C.c_double(), C.c_int64(), C.c_int64(), C.c_double()
    lib.snb_profile_end(C.byref(gms), C.byref(gl2), C.byref(tl2), C.byref(macs))
    snb_dist.barrier()
    if rank == 0:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak_tf, peak_hbm, which = peaks()
        gemm_ms_step = gms.value / nprof
        alg = ALG_FLOP_PER_TRAIN_RAY * B
        achieved = alg / (gemm_ms_step * 1e-3) / 1e12
        traffic, traffic_src = newest_ncu_traffic()
        roof = {"bound": "tensor", "kernel": "snb_chain_kernel + snb_gemm_kernel (tcgen05 GEMMs: chained MLP passes, wgrad, head rows)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": which, "frac_of_burst_peak": achieved / d["bf16_tflops"] if "bf16_tflops" in d else None,
                "per": "step: algorithmic MLP FLOPs of one step / summed GEMM launch time (one rank)",
                "note": "the GEMM time comes from 4 separately event-timed steps (launch-serialised): +-4 % against the timed region; "
                        "algorithmic FLOPs are the reference architecture's (SURVEY 8d) - the linear feats_from_xyz layer is folded "
                        "into the head first layers, so ~10 % fewer are executed (executed_tflops)",
                "gemm_launches_per_step": gl2.value // nprof, "gemm_ms_per_step": gemm_ms_step,
                "executed_tflops": 2 * macs.value / nprof / (gemm_ms_step * 1e-3) / 1e12,
                "step_basis_tflops": alg / (ms_total / args.steps * 1e-3) / 1e12,
                "gemm_share_of_step": gemm_ms_step / (ms_total / args.steps)}
    if rank == 0 and not args.no_extras:
        # secondary metric: no-grad render throughput (samples/s), chunked like batched_inference (no collectives)
        nr = 4 * 40960
        from semnerf_b200 import synth
        rr, ee = synth.make_rays(nr, seed=7)
        rr, ee = rr.to(dev), ee.to(dev)
        render = {}
        for name, rkeys, flop in (("main_and_solar_pass", ("rgb_coarse", "depth_coarse", "semantic_label_coarse", "sun_sc_coarse"),
                                   ALG_FLOP_PER_RENDER_SAMPLE),
                                  ("main_pass", ("rgb_coarse", "depth_coarse", "semantic_label_coarse"), 5_641_216)):
            tr.render_image(rr[:40960], ee[:40960], keys=rkeys)
            torch.cuda.synchronize()
            times = []
            for _ in range(3):   # one render of 4 chunks is ~0.1 s: a single allocator hiccup would halve the number
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                tr.render_image(rr, ee, keys=rkeys)
                r1.record()
                torch.cuda.synchronize()
                times.append(r0.elapsed_time(r1))
            rs = nr * N_SAMPLES / (min(times) * 1e-3)
            render[name] = {"samples_per_s": rs, "rays": nr, "chunk_rays": 40960, "tensor_frac": rs * flop / 1e12 / peak_tf,
                            "ms_of_3_renders": [round(t, 2) for t in times]}
        del rr, ee
        try:
            # at the training batch (8192 rays), the reference's render chunk (40 960 rays) and at 4 chunks
            hbm = {f"rays_{n}": hbm_kernel_rooflines(lib, dev, peak_hbm, n) for n in (8192, 40960, 163840)}
        except Exception as e:   # secondary numbers must not take the headline line down
            hbm = {"error": str(e)[:200]}
    if world == 1 and not args.no_extras:
        # ---- the reference's default batch (1024 rays / step, configs/pipelines/rs_semantic.toml) ----------------------
        sb = {k: v.to(dev) for k, v in make_batch(1024, seed=5).items()}
        for i in range(3):
            tr.training_step(sb, epoch=3)
        ms_b, _ = timed_steps(lambda i: tr.training_step(sb, epoch=3), 50, dev, world)
        small = {"rays_per_step": 1024, "value": 1024 * 50 / (ms_b * 1e-3), "unit": "rays/s", "ms_per_step": ms_b / 50,
                 "cuda_graph": tr.use_graph, "frac_of_headline": 1024 * 50 / (ms_b * 1e-3) / value}
        # ---- the early-training step: the depth-supervision batch of the first 25 % of the steps rides along
        # (semantic/components/training_step.py:31-49; its batch size is the rgb batch's: framework/pipelines.py:107-118)
        from semnerf_b200 import synth
        dr, de = synth.make_rays(B, seed=77)
        _, _, dd = synth.make_targets(dr, N_CLASSES, seed=77)
        db = {"rays": dr.to(dev), "extras": de.to(dev), "depths": dd.view(-1, 1).to(dev), "weights": torch.ones(B, 1, device=dev)}
        for i in range(3):
            tr.training_step(resident[0], epoch=3, depth_batch=db)
        ms_d, _ = timed_steps(lambda i: tr.training_step(resident[i % 4], epoch=3, depth_batch=db), 5, dev, world)
        early = {"workload": "the same step + the depth-supervision batch (trunk + sigma, DepthLoss) of the first 25 % of training",
                 "rgb_rays_per_step": B, "depth_rays_per_step": B, "value": B * 5 / (ms_d * 1e-3), "unit": "rgb rays/s",
                 "ms_per_step": ms_d / 5, "alg_tflops": (ALG_FLOP_PER_TRAIN_RAY + 64 * 11_320_320) * B * 5 / (ms_d * 1e-3) / 1e12}
        del db, sb
        tr._bufs.clear()
        tr._graphs.clear()
        torch.cuda.empty_cache()
        # ---- the "kernel to beat": the same step as eager cuBLAS + ATen on this GPU ------------------------------------
        eager = {"what": "the oracle's restatement of the reference path on CUDA tensors (eager cuBLAS + ATen, autograd, torch Adam): "
                         "TF32 'high' as in run/run_template.toml:15, and under torch.autocast(bfloat16)"}
        for n_e in (B, 1024):
            for mode in ("tf32", "bf16"):
                try:
                    eager[f"rays_{n_e}_{mode}"] = gpu_eager_steps(dev, n_e, mode)
                except Exception as e:
                    eager[f"rays_{n_e}_{mode}"] = None
                    eager[f"rays_{n_e}_{mode}_error"] = str(e)[:160]
                    torch.cuda.empty_cache()
        if eager.get(f"rays_{B}_tf32"):
            eager["speedup_vs_tf32"] = value / eager[f"rays_{B}_tf32"]
        if eager.get(f"rays_{B}_bf16"):
            eager["speedup_vs_bf16_autocast"] = value / eager[f"rays_{B}_bf16"]
        if eager.get("rays_1024_tf32") and small:
            eager["speedup_vs_tf32_at_1024"] = small["value"] / eager["rays_1024_tf32"]
        if not args.no_cpu:
            rps, cores, _ = cpu_training_steps(args.cpu_rays, 10, 1)   # ~10-15 s of CPU work
            cpu = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_rays} rays x {N_SAMPLES} samples per step (the GPU arm steps {B}; the metric is rays/s), "
                             "10 timed steps of the same training step"}
            cfg1 = cpu_config1()
    if rank == 0:
        line = {
            "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, graph=graph),
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "strong_scaling": strong, "batch_1024": small, "early_training_step": early,
            "gpu_eager_baseline": eager, "cpu_baseline_config1": cfg1, "render": render, "hbm_kernels": hbm,
            "loss": loss_host,
        }
        emit(line)
    snb_dist.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


_JSON_OUT = None


def emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner, build
    chatter) was redirected to stderr at start-up"""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8192, help="rays per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=512, help="rays per step of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step as one CUDA graph (single GPU default: on)")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline + e2e + roofline only (skip the strong-scaling, 1024-ray, early-step, eager, render, HBM legs)")
    ap.add_argument("--module-losses", action="store_true",
                    help="render_rays() + the reference-shaped loss modules instead of the fused K3 + loss kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

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
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = f"/tmp/snb_clocks_{os.getpid()}.csv"
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], None, set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


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


def workload_config(args, cpu=False):
    return {
        "workload": "Semantic-NeRF training step (BASELINE configs[1]): semantic 8x512 SIREN MLP, C=6, S=64, "
                    "solar-correction pass (sc_lambda=0.05), SatNerfLoss + SemanticLoss(ignore car) + "
                    "SemanticCarRegLoss, Adam; steady-state step (after the depth-supervision drop)",
        "rays_per_gpu_per_step": args.cpu_rays if cpu else args.batch, "samples_per_ray": N_SAMPLES,
        "n_classes": N_CLASSES, "parallelism": f"dp{args.gpus}",
        "losses": "loss modules on render_rays()" if getattr(args, "module_losses", False) else "fused into the compositing kernel",
        "l2": "per-step activation working set (~30 GB at 8192 rays) >> 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------------------
# HBM-bound kernels of the path (K1 sample + encode, K3 composite) timed alone against the measured copy peak
# ------------------------------------------------------------------------------------------------------
def hbm_kernel_rooflines(lib, dev, peak_hbm, n=40960):
    """Algorithmic bytes (DESIGN.md 4 / SURVEY 8d) / CUDA-event time, L2 flushed between repetitions."""
    import types
    from semnerf_b200 import synth
    from semnerf_b200._lib import check, ptr, stream
    from semnerf_b200.autograd import t_steps
    from semnerf_b200.model import RSSemanticNeRFB200
    from semnerf_b200.trainer import default_cfgs
    S, C = N_SAMPLES, N_CLASSES
    n_out, P = 9 + C, n * N_SAMPLES
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

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

    cfgs = default_cfgs("semantic", n_samples=S, sc_lambda=0.05)
    model = RSSemanticNeRFB200(cfgs, types.SimpleNamespace(semantic_n_classes=C)).to(dev)
    emb = torch.nn.Embedding(50, 4).to(dev)
    rays, extras = synth.make_rays(n, seed=1)
    rays, extras = rays.to(dev), extras.to(dev)
    zv = torch.empty(n, S, device=dev)
    enc = torch.empty(P, model.enc_ld, dtype=torch.bfloat16, device=dev)
    enc_sc = torch.empty_like(enc)
    aux = torch.empty(P, 16, dtype=torch.bfloat16, device=dev)
    ts, ew = t_steps(S, dev), emb.weight.detach().contiguous()
    k1_args = (ptr(rays), ptr(extras), None, 3, None, 0, ptr(ts), ptr(ew), 50, 4, None, None, None, None, 0, n, S, model.kind, 0,
               ptr(zv), ptr(enc), ptr(enc_sc), ptr(aux), None, stream())

    def k1x4():
        for _ in range(4):
            check(lib.snb_sample_encode(*k1_args), "k1")
    res = {}
    t = timed(k1x4) / 4
    b = n * 48 + P * (4 + 2 * model.enc_ld * 2 + 32)
    res["k1_sample_encode"] = {"bound": "hbm", "achieved": b / t / 1e9, "peak": peak_hbm, "unit": "GB/s",
                               "frac": b / t / 1e9 / peak_hbm, "bytes": b, "us": t * 1e6,
                               "note": "write-only kernel (main + solar rows); the measured write-only peak is ~3.9 TB/s"}
    out = torch.rand(P, n_out, device=dev)
    z = torch.sort(torch.rand(n, S, device=dev), dim=1).values
    rgb, depth = torch.empty(n, 3, device=dev), torch.empty(n, device=dev)
    w, T = torch.empty(n, S, device=dev), torch.empty(n, S, device=dev)
    sem, lab = torch.empty(n, C, device=dev), torch.empty(n, dtype=torch.int64, device=dev)
    t = timed(lambda: check(lib.snb_composite_forward(ptr(out), ptr(z), n, S, n_out, C, 0, ptr(rgb), ptr(depth), ptr(w), ptr(T),
                                                      ptr(sem), ptr(lab), stream()), "k3f"))
    b = n * (S * (4 * n_out + 4 + 8) + 12 + 4 + 4 * C + 8)
    res["k3_composite_forward"] = {"bound": "hbm", "achieved": b / t / 1e9, "peak": peak_hbm, "unit": "GB/s",
                                   "frac": b / t / 1e9 / peak_hbm, "bytes": b, "us": t * 1e6}
    g_rgb, g_d, g_w = torch.rand(n, 3, device=dev), torch.rand(n, device=dev), torch.rand(n, S, device=dev)
    g_sem, g_dir, g_out = torch.rand(n, C, device=dev), torch.rand(P, n_out, device=dev), torch.empty(P, n_out, device=dev)
    t = timed(lambda: check(lib.snb_composite_backward(ptr(out), ptr(z), n, S, n_out, C, 0, ptr(g_rgb), ptr(g_d), ptr(g_w), None,
                                                       ptr(g_sem), ptr(g_dir), ptr(g_out), stream()), "k3b"))
    b = n * (S * (4 * n_out + 4 + 4 + 4 * n_out + 4 * n_out) + 12 + 4 + 4 * C)
    res["k3_composite_backward"] = {"bound": "hbm", "achieved": b / t / 1e9, "peak": peak_hbm, "unit": "GB/s",
                                    "frac": b / t / 1e9 / peak_hbm, "bytes": b, "us": t * 1e6}
    res["rays"] = n
    return res


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
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
    B = args.batch
    cfgs = default_cfgs("semantic", n_samples=N_SAMPLES, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=0)
    tr = Trainer(cfgs, "semantic", N_CLASSES, device=dev, car_index=CAR_INDEX, world=world, rank=rank, seed=0,
                 fused_loss=not args.module_losses, graph=args.graph)
    host = [make_batch(B, seed=100 * rank + i, pinned=True) for i in range(4)]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def step_resident(i):
        return tr.training_step(resident[i % len(resident)], epoch=3, ray_offset=rank * B)

    for i in range(args.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    # ---- timed region 1: inputs resident in HBM, device-timed ----------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    lib.snb_profile_begin(0)
    snb_dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step_resident(i)
    e1.record()
    torch.cuda.synchronize()
    snb_dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_total = ms.item()
    clocks = sampler.stop() if sampler else None
    gl, tl = C.c_int64(), C.c_int64()
    lib.snb_profile_end(None, C.byref(gl), C.byref(tl), None)
    launches = tl.value
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- timed region 2: end to end through the public API, host buffers ------------------------------
    snb_dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        hb = host[i % len(host)]
        b = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
        loss = tr.training_step(b, epoch=3, ray_offset=rank * B)
        loss_host = loss.item()          # device -> host read of the step's result
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t_e2e, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * args.steps / t_e2e.item()
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    # ---- roofline of the dominant kernels (the tcgen05 GEMMs), timed live with CUDA events per launch ------
    # every rank runs these steps (they contain the gradient all-reduce); rank 0 reports its own launches
    roof = cpu = render = hbm = None
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
        peak_tf, peak_hbm, which = peaks()
        gemm_ms_step = gms.value / nprof
        alg = ALG_FLOP_PER_TRAIN_RAY * B
        achieved = alg / (gemm_ms_step * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "snb_chain_kernel + snb_gemm_kernel (tcgen05 GEMMs: chained MLP passes, wgrad, head rows)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": NCU_TRAFFIC_BYTES_PER_STEP,
                "peak_source": which, "per": "step: algorithmic MLP FLOPs of one step / summed GEMM launch time (one rank)",
                "gemm_launches_per_step": gl2.value // nprof, "gemm_ms_per_step": gemm_ms_step,
                "executed_tflops": 2 * macs.value / nprof / (gemm_ms_step * 1e-3) / 1e12,
                "gemm_share_of_step": gemm_ms_step / (ms_total / args.steps)}
        # secondary metric: no-grad render throughput (samples/s), chunked like batched_inference (no collectives)
        nr = 4 * 40960
        from semnerf_b200 import synth
        rr, ee = synth.make_rays(nr, seed=7)
        rr, ee = rr.to(dev), ee.to(dev)
        rkeys = ("rgb_coarse", "depth_coarse", "semantic_label_coarse", "sun_sc_coarse")   # sun_sc: keeps the solar pass in
        tr.render_image(rr[:40960], ee[:40960], keys=rkeys)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        tr.render_image(rr, ee, keys=rkeys)
        r1.record()
        torch.cuda.synchronize()
        rs = nr * N_SAMPLES / (r0.elapsed_time(r1) * 1e-3)
        render = {"samples_per_s": rs, "rays": nr, "chunk_rays": 40960, "passes": "main + solar-correction",
                  "tensor_frac": rs * ALG_FLOP_PER_RENDER_SAMPLE / 1e12 / peak_tf}
        try:
            # at the reference's render chunk (40 960 rays: 43-350 us launches, ramp and tail included) and at 4 chunks
            hbm = {"chunk_40960": hbm_kernel_rooflines(lib, dev, peak_hbm, 40960),
                   "rays_163840": hbm_kernel_rooflines(lib, dev, peak_hbm, 163840)}
        except Exception as e:   # secondary numbers must not take the headline line down
            hbm = {"error": str(e)[:200]}
        if world == 1 and not args.no_cpu:
            rps, cores, _ = cpu_training_steps(args.cpu_rays, 10, 1)   # ~10-15 s of CPU work
            cpu = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_rays} rays x {N_SAMPLES} samples, 10 timed steps of the same training step"}
    if rank == 0:
        line = {
            "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "render": render, "hbm_kernels": hbm, "loss": loss_host,
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
    ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph (single GPU)")
    ap.add_argument("--module-losses", action="store_true",
                    help="render_rays() + the reference-shaped loss modules instead of the fused K3 + loss kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

"""GPU diagnostics: run each kernel family against its reference and print error statistics.

    python tests/gpu_check_manual.py            # every group, each in its own subprocess
    python tests/gpu_check_manual.py gemm k3    # selected groups (in-process)

Used during bring-up through `gpurun`; the pytest `-m gpu` suite asserts the same comparisons.
Test infrastructure: imports oracle/.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm", "gemm_mn", "k1", "k3", "mlp", "render"]


def stats(name, got, ref, tol=None):
    import torch
    got, ref = got.double().flatten(), ref.double().flatten()
    err = (got - ref).abs()
    rel = err.max().item() / max(1e-30, ref.abs().max().item())
    cos = float((got @ ref) / max(1e-300, (got.norm() * ref.norm()).item()))
    bad = "" if tol is None or err.max().item() <= tol else "   <-- FAIL"
    nan = int(torch.isnan(got).sum().item())
    print(f"  {name:34s} max|d| {err.max().item():.3e}  mean|d| {err.mean().item():.3e}  rel {rel:.3e}  cos {cos:.8f}"
          f"  nan {nan}{bad}")
    return err.max().item()


def group_gemm():
    import torch
    from semnerf_b200 import _lib
    from semnerf_b200._lib import ptr, stream, check
    lib = _lib.load()
    dev = "cuda"
    torch.manual_seed(0)
    print("sms", lib.snb_device_sms())
    for (M, N, K) in [(128, 256, 64), (256, 512, 512), (1000, 512, 576), (4096, 256, 256), (300, 16, 1792), (8192, 1024, 576)]:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        B = (torch.randn(N, K, device=dev) * (1.0 / K ** 0.5)).bfloat16()
        bias = torch.randn(N, device=dev) * 0.1
        ref = A.float() @ B.float().t()
        print(f"[K-major] M={M} N={N} K={K}")
        if N >= 64:
            # linear
            out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_LINEAR, ptr(out), None, N, None,
                                    ptr(bias), 1.0, 1, stream()), "gemm linear")
            torch.cuda.synchronize()
            stats("linear(acc+bias)", out.float(), (ref + bias).bfloat16().float(), 0.05)
            # sin with derivative
            o0 = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            o1 = torch.zeros(M, max(N // 32, 1), device=dev, dtype=torch.int32)   # sign mask of the derivative
            check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_SIN, ptr(o0), ptr(o1), N, None,
                                    ptr(bias), 3.0, 1, stream()), "gemm sin")
            torch.cuda.synchronize()
            y = 3.0 * (ref + bias)
            stats("sin(w0*(acc+b))", o0.float(), torch.sin(y), 0.02)
            pos = torch.tensor([(c >> 1) + 16 * (c & 1) for c in range(32)], device=dev, dtype=torch.int32)
            bits = ((o1.view(M, -1, 1) >> pos) & 1).reshape(M, -1)[:, :N].bool()
            cy = torch.cos(y)
            print("   sign mask == [cos < 0] away from zero crossings:", bool(torch.equal(bits[cy.abs() > 1e-3], (cy < 0)[cy.abs() > 1e-3])))
            # mul
            mul = torch.randn(M, N, device=dev).bfloat16()
            o2 = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_MUL, ptr(o2), None, N, ptr(mul),
                                    None, 1.0, 1, stream()), "gemm mul")
            torch.cuda.synchronize()
            stats("acc*mul", o2.float(), ref * mul.float(), 0.08)
        else:
            o3 = torch.zeros(M, 16, device=dev, dtype=torch.float32)
            check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_F32ROWS, ptr(o3), None, 16, None,
                                    None, 1.0, 1, stream()), "gemm f32rows")
            torch.cuda.synchronize()
            stats("f32 rows (N=16)", o3, ref, 2e-3)


def group_gemm_mn():
    import torch
    from semnerf_b200 import _lib
    from semnerf_b200._lib import ptr, stream, check
    lib = _lib.load()
    dev = "cuda"
    torch.manual_seed(1)
    # wgrad form: G[Mf,Nf] = dY[P,Mf]^T X[P,Nf]
    for (P, Mf, Nf, splits) in [(64, 128, 256, 1), (512, 128, 256, 1), (4096, 512, 512, 4), (5000, 256, 64, 3),
                                (4096, 512, 16, 2), (70000, 1024, 512, 9)]:
        dY = (torch.randn(P, Mf, device=dev) * 0.1).bfloat16()
        X = torch.randn(P, Nf, device=dev).bfloat16()
        ref = dY.float().t() @ X.float()
        G = torch.zeros(Mf, Nf, device=dev, dtype=torch.float32)
        check(lib.snb_gemm_bf16(ptr(dY), Mf, ptr(X), Nf, Mf, Nf, P, 1, 1, _lib.EPI_WGRAD, ptr(G), None, Nf, None, None,
                                1.0, splits, stream()), "gemm wgrad")
        torch.cuda.synchronize()
        print(f"[MN-major wgrad] P={P} Mf={Mf} Nf={Nf} splits={splits}")
        stats("dY^T X", G, ref, 1e-2 * max(1.0, ref.abs().max().item()))
        # accumulate a second time: must double
        check(lib.snb_gemm_bf16(ptr(dY), Mf, ptr(X), Nf, Mf, Nf, P, 1, 1, _lib.EPI_WGRAD, ptr(G), None, Nf, None, None,
                                1.0, splits, stream()), "gemm wgrad")
        torch.cuda.synchronize()
        stats("accumulated twice", G, 2 * ref, 2e-2 * max(1.0, ref.abs().max().item()))


def group_k1():
    import torch
    from oracle import render_oracle as O
    from semnerf_b200.model import RSSemanticNeRFB200, SatNeRFB200
    from semnerf_b200.autograd import encode_rays
    from tests.helpers import make_cfgs
    dev = "cuda"
    for kind in ("semantic", "satnerf"):
        spec = O.ModelSpec(kind=kind, n_classes=6)
        for S in (2, 8, 64, 128):
            n = 200
            rays, extras = O.synthetic_rays(n, seed=S)
            u = torch.rand(n, S, generator=torch.Generator().manual_seed(S))
            params, emb = O.make_params(spec, seed=3)
            cfgs = make_cfgs(spec, S, 0.05)
            model = (RSSemanticNeRFB200(cfgs, type("D", (), {"semantic_n_classes": 6})()) if kind == "semantic"
                     else SatNeRFB200(cfgs)).to(dev)
            model.load_state_dict(params)
            z, enc, enc_sc, aux, sky = encode_rays(model, emb.to(dev), rays.to(dev), extras.to(dev), S, u=u.to(dev),
                                                   want_sc=True)
            torch.cuda.synchronize()
            z_ref = O.sample_z(rays, S, u)
            print(f"[k1] {kind} S={S}: z bit-exact: {torch.equal(z.cpu(), z_ref)}")
            stats("z_vals", z.cpu(), z_ref, 0.0)
            for nm, e, dirs in (("enc", enc, rays[:, 3:6]), ("enc_sc", enc_sc, extras[:, :3])):
                xyz = O.sample_points(rays[:, :3], dirs, z_ref).reshape(-1, 3)
                ref = O.posenc(xyz, 10) if kind == "semantic" else xyz
                k0 = spec.k0
                e = e.float().cpu()
                lo0 = k0 if kind == "semantic" else 2 * k0
                stats(nm + " hi+lo", e[:, :k0] + e[:, lo0:lo0 + k0], ref, 3e-5)
                print("   pad zero:", bool((e[:, lo0 + k0:] == 0).all()))
            a = aux.float().cpu().view(n, S, 16)
            t = emb[extras[:, 3].long()]
            ref_aux = torch.cat([torch.ones(n, 1), extras[:, :3], t, torch.zeros(n, 8)], 1)
            stats("aux", a[:, 0], ref_aux.bfloat16().float(), 0.0)
            stats("aux bcast", a[:, -1], a[:, 0], 0.0)
            sky_ref = torch.sigmoid(torch.relu(extras[:, :3] @ params["sky_color.0.weight"].t() + params["sky_color.0.bias"])
                                    @ params["sky_color.2.weight"].t() + params["sky_color.2.bias"])
            stats("sky", sky.cpu(), sky_ref, 1e-6)
        # philox path: in [0,1), deterministic, different per seed
        z1 = encode_rays(model, emb.to(dev), rays.to(dev), extras.to(dev), S, seed=7)[0]
        z2 = encode_rays(model, emb.to(dev), rays.to(dev), extras.to(dev), S, seed=7)[0]
        z3 = encode_rays(model, emb.to(dev), rays.to(dev), extras.to(dev), S, seed=8)[0]
        zl = O.sample_z(rays, S, torch.zeros(n, S))
        zh = O.sample_z(rays, S, torch.ones(n, S))
        print(f"[k1] philox: deterministic {torch.equal(z1, z2)}  seed-dependent {not torch.equal(z1, z3)}  "
              f"in-bin {bool(((z1.cpu() >= zl) & (z1.cpu() <= zh)).all())}  mean-u "
              f"{((z1.cpu() - zl) / (zh - zl).clamp_min(1e-9)).mean().item():.4f}")


def group_k3():
    import torch
    from oracle import render_oracle as O
    from semnerf_b200.autograd import Composite
    dev = "cuda"
    torch.manual_seed(0)
    for (n, S, C) in [(257, 64, 6), (64, 2, 6), (33, 8, 5), (40, 128, 6), (50, 64, 0), (19, 200, 3)]:
        n_out = 9 + C
        out = torch.rand(n, S, n_out)
        out[..., 3] = torch.rand(n, S) * 40 * (torch.rand(n, S) > 0.3)   # sigmas incl. exact zeros
        out[0, :, 3] = 0.0          # empty ray
        out[1, :, 3] = 1e4          # opaque at the first sample (T underflow)
        out[2, :, :3] = 3.0         # clamp active
        rays, _ = O.synthetic_rays(n, seed=S)
        z = O.sample_z(rays, S, torch.rand(n, S))
        o_ref = out.clone().double().requires_grad_(True)
        ref = O.composite(o_ref, z.double(), C)
        o_gpu = out.to(dev).requires_grad_(True)
        rgb, depth, w, T, sem, label = Composite.apply(o_gpu, z.to(dev), C)
        print(f"[k3] n={n} S={S} C={C}")
        stats("rgb", rgb.cpu(), ref["rgb"], 2e-5)
        stats("depth", depth.cpu(), ref["depth"], 2e-5)
        stats("weights", w.cpu(), ref["weights"], 2e-6)
        stats("transparency", T.cpu(), ref["transparency"], 2e-6)
        if C:
            stats("semantic_logits", sem.cpu(), ref["semantic_logits"], 2e-5)
            print("   label agreement:", (label.cpu() == ref["semantic_label"]).float().mean().item())
        # backward with every upstream gradient present
        g = torch.Generator().manual_seed(5)
        gr, gd, gw, gt = torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, S, generator=g), \
            torch.randn(n, S, generator=g)
        gs = torch.randn(n, max(C, 1), generator=g)[:, :C]
        loss_ref = (ref["rgb"] * gr).sum() + (ref["depth"] * gd).sum() + (ref["weights"] * gw).sum() + \
            (ref["transparency"] * gt).sum()
        loss = (rgb * gr.to(dev)).sum() + (depth * gd.to(dev)).sum() + (w * gw.to(dev)).sum() + (T * gt.to(dev)).sum()
        if C:
            loss_ref = loss_ref + (ref["semantic_logits"] * gs).sum()
            loss = loss + (sem * gs.to(dev)).sum()
        loss_ref.backward()
        loss.backward()
        stats("grad out (all upstream)", o_gpu.grad.cpu(), o_ref.grad, 5e-4 * max(1.0, o_ref.grad.abs().max().item()))


def _build(kind, C, dev, seed=1, trained_like=False, S=64, sc=0.05):
    import torch
    from oracle import render_oracle as O
    from semnerf_b200.model import RSSemanticNeRFB200, SatNeRFB200
    from tests.helpers import make_cfgs
    spec = O.ModelSpec(kind=kind, n_classes=C)
    params, emb = O.make_params(spec, seed=seed, trained_like=trained_like)
    cfgs = make_cfgs(spec, S, sc)
    model = (RSSemanticNeRFB200(cfgs, type("D", (), {"semantic_n_classes": C})()) if kind == "semantic"
             else SatNeRFB200(cfgs)).to(dev)
    model.load_state_dict(params)
    t = torch.nn.Embedding(spec.vocab, spec.tau).to(dev)
    t.weight.data.copy_(emb)
    return spec, params, emb, cfgs, model, t


def group_mlp():
    import torch
    from oracle import render_oracle as O
    dev = "cuda"
    for kind, C in (("semantic", 6), ("satnerf", 0)):
        spec, params, emb, cfgs, model, t = _build(kind, C, dev)
        P = 1000
        g = torch.Generator().manual_seed(0)
        xyz = torch.rand(P, 3, generator=g) * 2 - 1
        sun = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1)
        tt = torch.randn(P, 4, generator=g)
        p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
        tt64 = tt.double().requires_grad_(True)
        ref, hidden, f = O.mlp_forward(p64, spec, xyz.double(), sun.double(), tt64, return_hidden=True)
        tg = tt.to(dev).requires_grad_(True)
        out = model(xyz.to(dev), input_sun_dir=sun.to(dev), input_t=tg)
        torch.cuda.synchronize()
        print(f"[mlp] {kind}: forward (P={P})")
        names = ["rgb0", "rgb1", "rgb2", "sigma", "sun", "sky0", "sky1", "sky2", "beta"] + [f"sem{c}" for c in range(C)]
        for j, nm in enumerate(names):
            stats(nm, out[:, j].detach().cpu(), ref[:, j].detach())
        w = torch.randn(ref.shape, generator=g).double()
        (ref * w).sum().backward()
        (out * w.to(dev).float()).sum().backward()
        torch.cuda.synchronize()
        print(f"[mlp] {kind}: gradients")
        grads = model.named_grads()
        num = da = db = 0.0
        for k in p64:
            ga, gb = grads[k].double().cpu().flatten(), p64[k].grad.flatten()
            c = float(ga @ gb) / max(1e-300, float(ga.norm() * gb.norm()))
            print(f"  {k:32s} cos {c:.6f}  |ref| {gb.norm().item():.3e}  |got| {ga.norm().item():.3e}")
            num += float(ga @ gb); da += float(ga @ ga); db += float(gb @ gb)
        print(f"  GLOBAL cosine {num / (da * db) ** 0.5:.6f}")
        stats("grad input_t", tg.grad.cpu(), tt64.grad)


def group_render():
    import torch
    from oracle import render_oracle as O
    from semnerf_b200.renderer import B200Renderer
    dev = "cuda"
    for kind, C, trained in (("semantic", 6, False), ("semantic", 6, True), ("satnerf", 0, False)):
        S, n = 64, 512
        spec, params, emb, cfgs, model, t = _build(kind, C, dev, seed=7 if trained else 1, trained_like=trained, S=S)
        rays, extras = O.synthetic_rays(n, seed=11)
        u = torch.rand(n, S, generator=torch.Generator().manual_seed(3))
        p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        e64 = emb.clone().requires_grad_(True)
        ref = O.render_rays(p64, e64, spec, rays, extras, S, u=u, sc_lambda=0.05)
        renderer = B200Renderer(cfgs)
        models = {"coarse": model, "t": t}
        res = renderer.render_rays(models, rays.to(dev), extras.to(dev), render_options={"u": u.to(dev)})
        torch.cuda.synchronize()
        print(f"[render] {kind} trained_like={trained}: keys {sorted(res.keys())}")
        for k, v in ref.items():
            if k.startswith("_"):
                continue
            if v.dtype.is_floating_point:
                stats(k, res[k].detach().cpu(), v.detach())
            else:
                print(f"  {k:34s} agreement {(res[k].cpu() == v).float().mean().item():.5f}")
        print(f"  PSNR(ours, oracle) = {O.psnr(res['rgb_coarse'].detach().cpu(), ref['rgb_coarse'].detach()):.2f} dB")
        gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(9))
        loss_ref = O.satnerf_loss(ref, gt)
        lab = torch.randint(0, max(C, 1), (n,), generator=torch.Generator().manual_seed(4))
        if C:
            loss_ref = loss_ref + O.semantic_loss(ref, lab)
        loss_ref.backward()
        loss = O.satnerf_loss(res, gt.to(dev))
        if C:
            loss = loss + O.semantic_loss(res, lab.to(dev))
        loss.backward()
        torch.cuda.synchronize()
        print(f"  loss ref {loss_ref.item():.6f} ours {loss.item():.6f}")
        grads = model.named_grads()
        num = da = db = 0.0
        for k in p64:
            ga, gb = grads[k].double().cpu().flatten(), p64[k].grad.double().flatten()
            c = float(ga @ gb) / max(1e-300, float(ga.norm() * gb.norm()))
            print(f"  {k:32s} cos {c:.6f}  |ref| {gb.norm().item():.3e}  |got| {ga.norm().item():.3e}")
            num += float(ga @ gb); da += float(ga @ ga); db += float(gb @ gb)
        print(f"  GLOBAL gradient cosine {num / (da * db) ** 0.5:.6f}")
        stats("grad embedding", t.weight.grad.cpu(), e64.grad)


if __name__ == "__main__":
    from semnerf_b200 import build as _build_lib
    _build_lib.build()
    sel = [a for a in sys.argv[1:] if not a.startswith("-")]
    if sel:
        for gname in sel:
            t0 = time.time()
            globals()["group_" + gname]()
            print(f"== {gname} done in {time.time() - t0:.1f}s", flush=True)
    else:
        rc = 0
        for gname in GROUPS:
            print(f"================ {gname} ================", flush=True)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), gname], cwd=ROOT, timeout=600)
            if r.returncode != 0:
                print(f"!! group {gname} exited with {r.returncode}", flush=True)
                rc = 1
        sys.exit(rc)

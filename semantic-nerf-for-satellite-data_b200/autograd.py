"""torch.autograd.Function wrappers around the C ABI (K1 encode, K2 MLP, K3 composite).

These replace the autograd graph PyTorch would build for the reference's eager implementation
(SURVEY 8a row a10).  All tensors stay PyTorch-owned; the kernels run on the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import HEADS_ALL, HEADS_DEPTH, HEADS_SOLAR, check, ptr, stream

_T_STEPS = {}


def t_steps(n_samples: int, device) -> torch.Tensor:
    """linspace(0,1,S) computed by the host exactly as the reference does (rendering.py:95)."""
    key = (n_samples, str(device))
    if key not in _T_STEPS:
        _T_STEPS[key] = torch.linspace(0, 1, n_samples).to(device)
    return _T_STEPS[key]


def _f32c(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def encode_rays(model, emb_weight, rays, extras, n_samples, u=None, z=None, seed=0, ray_offset=0, want_sc=False,
                want_sky=True, seed_dev=None):
    """K1: z_vals (N,S), enc / enc_sc (P, enc_ld) bf16, aux (P,16) bf16, sky (N,3).  No autograd here:
    the embedding gradient comes back through the MLP backward's aux gradient."""
    lib = _lib.load()
    rays, extras = _f32c(rays), _f32c(extras)
    n = rays.shape[0]
    dev = rays.device
    P = n * n_samples
    z_given = z is not None
    z_vals = _f32c(z).clone() if z_given else torch.empty(n, n_samples, dtype=torch.float32, device=dev)
    enc = torch.empty(P, model.enc_ld, dtype=torch.bfloat16, device=dev)
    enc_sc = torch.empty(P, model.enc_ld, dtype=torch.bfloat16, device=dev) if want_sc else None
    aux = torch.empty(P, 16, dtype=torch.bfloat16, device=dev)
    sky = torch.empty(n, 3, dtype=torch.float32, device=dev) if want_sky else None
    nerf = model.kind == _lib.MODEL_NERF      # no sky head; xyz is encoded like the semantic model's
    if nerf:
        sky, (w1, b1, w2, b2), hidden = None, (None, None, None, None), 0
    else:
        w1, b1, w2, b2 = model.sky_params()
        hidden = w1.shape[0]
    ew = _f32c(emb_weight.detach()) if emb_weight is not None else None
    check(lib.snb_sample_encode(ptr(rays), ptr(extras), ptr(_f32c(u)), seed, ptr(seed_dev), ray_offset, ptr(t_steps(n_samples, dev)),
                                ptr(ew), ew.shape[0] if ew is not None else 0, ew.shape[1] if ew is not None else 0,
                                ptr(w1), ptr(b1), ptr(w2), ptr(b2), hidden, n, n_samples,
                                _lib.K1_KIND[model.kind],
                                1 if z_given else 0, ptr(z_vals), ptr(enc), ptr(enc_sc), ptr(aux), ptr(sky), stream()),
          "snb_sample_encode")
    if nerf:   # the aux row of NeRF: [1, Mapping(4, 3)(view direction), 0...] (32 columns) instead of [1, sun_d, t]
        aux = nerf_aux(rays[:, 3:6], n_samples)
    return z_vals, enc, enc_sc, aux, sky


def nerf_aux(dirs, n_samples: int):
    """(N,3) view directions -> aux (N * n_samples, 32) bf16 (snb_nerf_aux)."""
    lib = _lib.load()
    dirs = _f32c(dirs)
    n = dirs.shape[0]
    aux = torch.empty(n * n_samples, 32, dtype=torch.bfloat16, device=dirs.device)
    check(lib.snb_nerf_aux(ptr(dirs), 3, n, n_samples, ptr(aux), stream()), "snb_nerf_aux")
    return aux


def _workspace(model, P, train, device):
    lib = _lib.load()
    nbytes = lib.snb_mlp_workspace_bytes(model._h, P, 1 if train else 0)
    if train:
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    cache = model.__dict__.setdefault("_ws_infer", {})
    key = (P, str(device))
    if key not in cache:
        cache.clear()
        cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return cache[key]


class MLPRays(torch.autograd.Function):
    """K2 on ray samples: (flat params, embedding table) -> packed (P, n_out) head outputs."""

    @staticmethod
    def forward(ctx, flat, emb_weight, model, enc, aux, sky, extras, n_rays, n_samples, head_mask, train):
        # `train` comes from the caller (mlp_rays): grad mode is always off inside forward(), and needs_input_grad is
        # True for a parameter even under torch.no_grad(), which would make every render save its activations
        lib = _lib.load()
        P = n_rays * n_samples
        ws = _workspace(model, P, train, enc.device)
        out = torch.empty(P, model.n_out_kernel, dtype=torch.float32, device=enc.device)
        packed = model.packed()
        check(lib.snb_mlp_forward(model._h, ptr(packed), ptr(ws), ws.numel(), P, ptr(enc), ptr(aux), ptr(sky),
                                  n_samples, head_mask, 1 if train else 0, ptr(out), stream()), "snb_mlp_forward")
        ctx.model, ctx.head_mask, ctx.dims = model, head_mask, (n_rays, n_samples)
        ctx.has_emb = emb_weight is not None
        if train:
            ctx.save_for_backward(flat, packed, ws, enc, aux, sky, extras, out,
                                  emb_weight if emb_weight is not None else flat.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        model, head_mask = ctx.model, ctx.head_mask
        n_rays, S = ctx.dims
        P = n_rays * S
        flat, packed, ws, enc, aux, sky, extras, out, emb_weight = ctx.saved_tensors
        g_out = _f32c(g_out)
        g_flat = torch.zeros_like(flat)
        want_emb = ctx.has_emb and head_mask == HEADS_ALL
        g_aux = torch.empty(P, 16, dtype=torch.float32, device=enc.device) if want_emb else None
        check(lib.snb_mlp_backward(model._h, ptr(packed), ptr(ws), ws.numel(), P, ptr(enc), ptr(aux), ptr(out),
                                   ptr(g_out), head_mask, ptr(g_flat), ptr(g_aux), None, stream()), "snb_mlp_backward")
        g_emb = torch.zeros_like(emb_weight) if want_emb else None
        sky_arg = sky if (head_mask == HEADS_ALL and model.kind != _lib.MODEL_NERF) else None
        if sky_arg is not None or g_aux is not None:
            vocab, tau = (emb_weight.shape if want_emb else (1, 0))
            check(lib.snb_ray_param_backward(model._h, ptr(flat.detach()), ptr(extras), ptr(sky_arg), ptr(g_out),
                                             ptr(g_aux), n_rays, S, model.n_out_kernel, tau, vocab,
                                             ptr(g_flat), ptr(g_emb), stream()), "snb_ray_param_backward")
        return g_flat, g_emb, None, None, None, None, None, None, None, None, None


class MLPPoints(torch.autograd.Function):
    """K2 on caller-supplied points: the reference's Model.forward (satnerf.py:208, rs_semantic.py:260)."""

    @staticmethod
    def forward(ctx, flat, t, model, xyz, sun_d, train):
        lib = _lib.load()
        xyz, sun_d, tt = _f32c(xyz), _f32c(sun_d), _f32c(t.detach())
        P = xyz.shape[0]
        dev = xyz.device
        enc = torch.empty(P, model.enc_ld, dtype=torch.bfloat16, device=dev)
        aux = torch.empty(P, 16, dtype=torch.bfloat16, device=dev)
        nerf = model.kind == _lib.MODEL_NERF   # `sun_d` is the view direction there; no sky head, no embedding
        if nerf:
            sky, (w1, b1, w2, b2), hidden = None, (None, None, None, None), 0
        else:
            sky = torch.empty(P, 3, dtype=torch.float32, device=dev)
            w1, b1, w2, b2 = model.sky_params()
            hidden = w1.shape[0]
        check(lib.snb_encode_points(ptr(xyz), ptr(sun_d), ptr(tt), tt.shape[1], ptr(w1), ptr(b1), ptr(w2), ptr(b2),
                                    hidden, P, _lib.K1_KIND[model.kind], ptr(enc), ptr(aux), ptr(sky),
                                    stream()), "snb_encode_points")
        if nerf:
            aux = nerf_aux(sun_d, 1)
        ws = _workspace(model, P, train, dev)
        out = torch.empty(P, model.n_out_kernel, dtype=torch.float32, device=dev)
        packed = model.packed()
        check(lib.snb_mlp_forward(model._h, ptr(packed), ptr(ws), ws.numel(), P, ptr(enc), ptr(aux), ptr(sky), 0,
                                  HEADS_ALL, 1 if train else 0, ptr(out), stream()), "snb_mlp_forward")
        ctx.model = model
        if train:
            ctx.save_for_backward(flat, packed, ws, enc, aux, sky, sun_d, out)
            ctx.tau = tt.shape[1]
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        model = ctx.model
        flat, packed, ws, enc, aux, sky, sun_d, out = ctx.saved_tensors
        P = out.shape[0]
        g_out = _f32c(g_out)
        g_flat = torch.zeros_like(flat)
        g_aux = torch.empty(P, 16, dtype=torch.float32, device=enc.device)
        check(lib.snb_mlp_backward(model._h, ptr(packed), ptr(ws), ws.numel(), P, ptr(enc), ptr(aux), ptr(out),
                                   ptr(g_out), HEADS_ALL, ptr(g_flat), ptr(g_aux), None, stream()), "snb_mlp_backward")
        if model.kind == _lib.MODEL_NERF:   # no sky_color parameters, no embedding
            return g_flat, None, None, None, None, None
        extras = torch.cat([sun_d, torch.zeros(P, 1, device=sun_d.device)], 1).contiguous()
        check(lib.snb_ray_param_backward(model._h, ptr(flat.detach()), ptr(extras), ptr(sky), ptr(g_out), None, P, 1,
                                         model.n_out_kernel, 0, 1, ptr(g_flat), None, stream()),
              "snb_ray_param_backward")
        return g_flat, g_aux[:, 4:4 + ctx.tau].contiguous(), None, None, None, None


def _wants_grad(*tensors) -> bool:
    """Does this call have to save activations for a backward pass?  (decided OUTSIDE the autograd.Function)"""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def mlp_points(model, xyz, sun_d, t):
    return MLPPoints.apply(model.flat, t, model, xyz, sun_d, _wants_grad(model.flat, t))


def mlp_rays(model, emb_weight, enc, aux, sky, extras, n_rays, n_samples, head_mask):
    """K2 on ray samples; inference (no saved activations, per-SM-pair L2 scratch) unless a gradient is required."""
    return MLPRays.apply(model.flat, emb_weight, model, enc, aux, sky, extras, n_rays, n_samples, head_mask,
                         _wants_grad(model.flat, emb_weight))


def mlp_fp32(model, xyz, sun_d, t, sky, rows_per_ray: int, head_mask: int = HEADS_ALL):
    """fp32 verification mode of K2 (snb_mlp_forward_fp32): fp32 weights straight from the flat parameter, fp32 FMA
    accumulation, no bf16 anywhere.  Inference only - there is no backward for this mode.
    xyz (P,3); sun_d / t / sky one row per ray when rows_per_ray > 1, else per point."""
    lib = _lib.load()
    if torch.is_grad_enabled() and model.flat.requires_grad:
        raise _lib.SnbError("fp32 mode is an inference / verification mode: call it under torch.no_grad()")
    xyz, sun_d, t, sky = _f32c(xyz), _f32c(sun_d), _f32c(t), _f32c(sky)
    P = xyz.shape[0]
    cache = model.__dict__.setdefault("_ws_fp32", {})
    nbytes = lib.snb_mlp_fp32_workspace_bytes(model._h, P)
    key = (nbytes, str(xyz.device))
    if key not in cache:
        cache.clear()
        cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=xyz.device)
    ws = cache[key]
    out = torch.empty(P, model.n_out_kernel, dtype=torch.float32, device=xyz.device)
    check(lib.snb_mlp_forward_fp32(model._h, ptr(model.flat.detach()), ptr(ws), ws.numel(), P, ptr(xyz), ptr(sun_d), ptr(t),
                                   ptr(sky), rows_per_ray, head_mask, ptr(out), stream()), "snb_mlp_forward_fp32")
    return out


class Composite(torch.autograd.Function):
    """K3: packed head outputs (N,S,n_out) + z_vals (N,S) -> rgb, depth, weights, transparency,
    semantic scores, label  (framework/util/rendering.py:4-34 + the tail of `inference`)."""

    @staticmethod
    def forward(ctx, out, z_vals, n_classes, flags=0):
        lib = _lib.load()
        out, z_vals = _f32c(out), _f32c(z_vals)
        n, s, n_out = out.shape
        dev = out.device
        rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        weights = torch.empty(n, s, dtype=torch.float32, device=dev)
        transp = torch.empty(n, s, dtype=torch.float32, device=dev)
        sem = torch.empty(n, n_classes, dtype=torch.float32, device=dev)
        label = torch.empty(n, dtype=torch.int64, device=dev)
        check(lib.snb_composite_forward(ptr(out), ptr(z_vals), n, s, n_out, n_classes, flags, ptr(rgb), ptr(depth),
                                        ptr(weights), ptr(transp), ptr(sem) if n_classes else None,
                                        ptr(label) if n_classes else None, stream()), "snb_composite_forward")
        ctx.save_for_backward(out, z_vals)
        ctx.n_classes, ctx.flags = n_classes, flags
        ctx.mark_non_differentiable(label)
        return rgb, depth, weights, transp, sem, label

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_w, g_t, g_sem, _g_label):
        lib = _lib.load()
        out, z_vals = ctx.saved_tensors
        n, s, n_out = out.shape
        g_out = torch.empty_like(out)
        check(lib.snb_composite_backward(ptr(out), ptr(z_vals), n, s, n_out, ctx.n_classes, ctx.flags, ptr(_f32c(g_rgb)),
                                         ptr(_f32c(g_depth)), ptr(_f32c(g_w)), ptr(_f32c(g_t)),
                                         ptr(_f32c(g_sem)) if ctx.n_classes else None, None, ptr(g_out), stream()),
              "snb_composite_backward")
        return g_out, None, None, None


LOSS_TERMS = ("color", "logbeta", "semantic", "car_reg", "sc_term2", "sc_term3", "ds", "semantic_logbeta")


def label_counts(labels, ray_mask, n_classes: int, ignore_index: int, car_label: int, counts=None):
    """snb_label_counts: the masked-mean denominators of the semantic losses as a device float[8]
    [rays in the CE mean, rays in the car term, out-of-range labels, 0, then the two statistics of the uncertainty-weighted
    semantic loss's pre-pass (CE sum, sum 1 / (2 beta^2)), 0, 0] - no host sync.  `labels` int64 (N), `ray_mask` uint8 (N) or
    None."""
    lib = _lib.load()
    if counts is None:
        counts = torch.zeros(8, dtype=torch.float32, device=labels.device)
    check(lib.snb_label_counts(ptr(labels), ptr(ray_mask), labels.numel(), n_classes, ignore_index, car_label, ptr(counts),
                               stream()), "snb_label_counts")
    return counts


def as_labels(semantic) -> torch.Tensor:
    """labels as the kernels read them: contiguous int64 (N).  The reference's dataset yields uint8
    (framework/util/img_utils.py::load_tensor_from_cls_geotiff; semantic/components/metrics.py:47)."""
    lab = semantic.reshape(-1)
    if lab.dtype != torch.int64:
        lab = lab.to(torch.int64)
    return lab.contiguous()


def as_ray_mask(mask):
    """semantic_sparsity_mask (bool, (N,) or (N,1)) -> contiguous uint8 (N), None stays None"""
    if mask is None:
        return None
    return mask.reshape(-1).to(torch.uint8).contiguous()


class CompositeLoss(torch.autograd.Function):
    """K3 + losses fused (snb_composite_loss): packed head outputs (N,S,n_out) + z_vals -> the scalar sum of the loss
    terms of one pass; the gradient of the packed rows is produced by the same kernel and handed to the MLP backward.
    `terms` (8,) is accumulated in place (logging; not differentiable)."""

    @staticmethod
    def forward(ctx, out, z_vals, n_classes, params, gt_rgb, labels, depth_gt, depth_w, counts, terms, ray_mask=None,
                reduce_stats=None):
        lib = _lib.load()
        out, z_vals = _f32c(out), _f32c(z_vals)
        n, s, n_out = out.shape
        g_out = torch.empty_like(out)
        mine = torch.zeros(8, dtype=torch.float32, device=out.device)
        if params.mode == 0 and params.sem_unc:
            # the uncertainty-weighted semantic loss is a product of two batch means: statistics pre-pass first (mode 3)
            pre = _lib.LossParams.from_buffer_copy(params)
            pre.mode = 3
            check(lib.snb_composite_loss(ptr(out), ptr(z_vals), n, s, n_out, n_classes, None, ptr(labels), ptr(ray_mask), None,
                                         None, ptr(counts), C.addressof(pre), None, counts[4:].data_ptr(), stream()),
                  "snb_composite_loss (statistics pre-pass)")
            if reduce_stats is not None:
                reduce_stats(counts[4:6])
        check(lib.snb_composite_loss(ptr(out), ptr(z_vals), n, s, n_out, n_classes, ptr(_f32c(gt_rgb)), ptr(labels),
                                     ptr(ray_mask), ptr(_f32c(depth_gt)), ptr(_f32c(depth_w)), ptr(counts), C.addressof(params),
                                     ptr(g_out), ptr(mine), stream()), "snb_composite_loss")
        ctx.save_for_backward(g_out)
        if terms is not None:
            terms.add_(mine)
        loss = mine.sum()
        if params.mode == 0 and params.color == 1:
            # (3 + mean log beta) / 2: the constant of the log-beta term (loss.py:26); a data-parallel shard carries its share
            loss = loss + 1.5 * n * params.inv_n
        if params.mode == 0 and params.sem_unc and (params.flags & _lib.COMPOSITE_BETA_S):
            loss = loss + 1.5 * params.lambda_s * n * params.inv_n   # the separate semantic uncertainty's log term (loss.py:27-30)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        (g_out,) = ctx.saved_tensors
        return g_out * g_loss, None, None, None, None, None, None, None, None, None, None, None

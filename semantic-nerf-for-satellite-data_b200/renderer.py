"""Renderers mirroring the reference's components interface.

``B200Renderer.render_rays(models, rays, extras, epoch, progress, render_options)`` has the
signature and the output dictionary (``*_coarse`` keys) of ``BaseRenderer.render_rays``
(framework/components/rendering.py:125-157) combined with ``SatNeRFRendering._model_rendering``
(baseline/components/rendering.py:12-67) / ``RSSemanticRendering._model_rendering``
(semantic/components/rendering.py:18-80), so ``BaseRayPipeline.forward``
(baseline/pipelines/base_ray_pipeline.py:34-52) and ``batched_inference``
(eval/utils/util.py:13-42) can call it unchanged.

render_options (all optional, default = reference behaviour):
  "u" (N,S) / "z_vals" (N,S): replace the random jitter (parity tests; the reference's
      ``given_z_vals`` hook, rendering.py:91-92);  "seed", "ray_offset": Philox key / global ray index;
  "heads": "all" | "depth" - "depth" evaluates only trunk + sigma (what the depth-supervision
      batch consumes, semantic/components/training_step.py:32-46) and skips the solar pass;
  "precision": "bf16" (default: the tcgen05 path) | "fp32" - the fp32 verification mode (fp32 weights and
      accumulation on the CUDA cores, inference only; also selected by cfgs.pipeline.precision = "fp32");
  "solar_pass": False - skip the solar-correction pass (its `*_sc` outputs feed only the training loss).
"""
from __future__ import annotations

import torch

from ._lib import COMPOSITE_BETA_S, COMPOSITE_NO_CLAMP, HEADS_ALL, HEADS_DEPTH, HEADS_SOLAR, MODEL_SEMANTIC
from . import _lib
from .autograd import (Composite, CompositeLoss, as_labels, as_ray_mask, encode_rays, label_counts, mlp_fp32,
                       mlp_rays)


class B200Renderer:
    def __init__(self, cfgs):
        self.cfgs = cfgs
        self.N_samples = cfgs.pipeline.n_samples
        self._calls = 0

    def render_rays(self, models: dict, rays, extras, epoch=None, progress=1.0, render_options={}):
        res = self._model_rendering(models, "coarse", self.cfgs, rays, extras, None, None, None, epoch=epoch,
                                    progress=progress, render_options=render_options)
        return {f"{k}_coarse": v for k, v in res.items()}

    def _model_rendering(self, models, typ, cfgs, rays, extras, xyz, z_vals, rays_d, epoch=None, progress=1.0,
                         render_options=None):
        opts = render_options or {}
        model = models[typ]
        emb = self._embedding(models, model)
        n, S = rays.shape[0], self.N_samples
        if z_vals is None:
            z_vals = opts.get("z_vals")
        depth_only = opts.get("heads", "all") == "depth"
        nerf = getattr(model, "variant", None) == "nerf"   # NeRFRendering (baseline/components/rendering.py:103-118)
        # "solar_pass": False skips the solar-correction pass (its outputs feed only the training loss; an evaluation render
        # that wants rgb / depth / labels need not pay for it - the reference always runs it when sc_lambda > 0)
        # (nerf.toml has no sc_lambda at all: NeRF has no sun head and no solar-correction pass)
        sc = getattr(cfgs.pipeline, "sc_lambda", 0.0) > 0 and not depth_only and bool(opts.get("solar_pass", True)) \
            and not nerf
        if n == 0:   # an empty ray batch renders to empty tensors, as the reference's eager code does
            return self._empty_result(model, rays, S, sc)
        self._calls += 1
        z, enc, enc_sc, aux, sky = encode_rays(model, emb, rays, extras, S, u=opts.get("u"), z=z_vals,
                                               seed=int(opts.get("seed", self._calls)),
                                               ray_offset=int(opts.get("ray_offset", 0)), want_sc=sc)
        C = model.semantic_n_classes
        mask = HEADS_DEPTH if depth_only else HEADS_ALL
        fp32 = opts.get("precision", getattr(cfgs.pipeline, "precision", "bf16")) == "fp32"
        if fp32:
            # sample positions exactly as the reference forms them (framework/components/rendering.py:113-115) and the
            # per-ray inputs in fp32; K1 above still supplies z_vals (bit-exact) and the per-ray sky colour (fp32)
            o, d, sun_d = rays[:, 0:3].float(), rays[:, 3:6].float(), extras[:, 0:3].float()
            t_ray = emb.detach()[extras[:, 3].long()].float() if emb is not None else torch.zeros(n, 4, device=rays.device)
            xyz_main = (o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)).reshape(-1, 3)
            if nerf:   # the per-ray input is the ENCODED view direction; the sun column is pinned to 1 (no lighting model)
                from .model import posenc_dirs
                out = mlp_fp32(model, xyz_main, posenc_dirs(d), None, None, S, mask)
                out[:, 4] = 1.0
                out = out.view(n, S, -1)
            else:
                out = mlp_fp32(model, xyz_main, sun_d, t_ray, sky, S, mask).view(n, S, -1)
        else:
            out = mlp_rays(model, emb, enc, aux, sky, extras, n, S, mask).view(n, S, -1)
        # NeRF's inference returns the raw composited colour (nerf.py:73-86); the other models clamp it to [0, 1]
        bs = getattr(model, "beta_s", 0)   # use_separate_beta_for_s: column 9 = beta_semantic, classes from column 10
        rgb, depth, weights, transp, sem, label = Composite.apply(
            out, z, C, (COMPOSITE_NO_CLAMP if nerf else 0) | (COMPOSITE_BETA_S if bs else 0))
        result = {
            "rgb": rgb, "depth": depth, "weights": weights, "transparency": transp,
            "albedo": out[..., :3], "sun": out[..., 4:5], "sky": out[..., 5:8], "beta": out[..., 8:9],
            "sigmas": out[..., 3],
        }
        if model.kind == MODEL_SEMANTIC:
            result["semantic_logits"] = sem
            result["semantic_label"] = label
            if bs:
                result["beta_semantic"] = out[..., 9:10]       # rs_semantic.py:90-96,126-127
        if getattr(model, "variant", None) == "snerf":   # snerf.py:86-96 returns neither beta nor sigmas
            del result["beta"], result["sigmas"]
        if nerf:                                         # nerf.py:80-86: rgb, depth, weights, transparency only
            result = {k: result[k] for k in ("rgb", "depth", "weights", "transparency")}
        if sc:
            # solar correction: second pass on o + sun_d*z, keeping weights / transparency / sun
            if fp32:
                xyz_sc = (o.unsqueeze(1) + sun_d.unsqueeze(1) * z.unsqueeze(2)).reshape(-1, 3)
                out_sc = mlp_fp32(model, xyz_sc, sun_d, None, None, S, HEADS_SOLAR).view(n, S, -1)
            else:
                out_sc = mlp_rays(model, emb, enc_sc, aux, None, extras, n, S, HEADS_SOLAR).view(n, S, -1)
            _, _, w_sc, t_sc, _, _ = Composite.apply(out_sc, z, 0)  # no semantic columns to composite
            result["weights_sc"] = w_sc
            result["transparency_sc"] = t_sc
            result["sun_sc"] = out_sc[..., 4:5]
        result["_z_vals"] = z
        return result


    def render_loss(self, models: dict, rays, extras, rgbs, semantic=None, *, color: str = "satnerf", lambda_s: float = 0.0,
                    ignore_index: int = -100, lambda_c: float = 0.0, car_label: int = -1, beta_min: float = 0.05,
                    depth=None, depth_weights=None, lambda_ds: float = 0.0, ignore_mask=None, global_rays=None,
                    reduce_counts=None, semantic_uncertainty: int = 0, render_options=None):
        """Training fast path: render + the losses that sit on the render outputs + their gradients
        with the compositing fused (snb_composite_loss), without materialising any per-sample output tensor.
        Equivalent to render_rays() followed by SNerfLoss / SatNerfLoss (color = "snerf" / "satnerf", with the solar
        correction when cfgs.pipeline.sc_lambda > 0; NeRF: NerfLoss = color "snerf" without a solar pass) [+ SemanticLoss +
        SemanticCarRegLoss when `semantic` labels are given, `ignore_mask` = the reference's semantic_sparsity_mask,
        semantic/components/training_step.py:58-88], or - with `depth` targets - to the depth-supervision pass + DepthLoss.
        semantic_uncertainty: 0 = SemanticLoss; 1 = SemanticUncertaintyLoss (`use_beta_for_s`: the cross-entropy mean weighted by
        the batch mean of 1 / (2 beta^2), semantic/components/loss.py:6-32,68-114); 2 = the same with beta detached.
        Data parallel: `global_rays` = rays of the GLOBAL batch (the means of the reference losses run over it) and
        `reduce_counts` = a callable that sum-all-reduces the masked-mean denominators in place; the per-rank losses then
        add up to the single-process loss on the concatenated batch, and so do the gradients.
        Returns (loss, terms) where terms is a (8,) tensor in the order of autograd.LOSS_TERMS (the log-beta entry without
        its constant 3/2)."""
        opts = render_options or {}
        model = models["coarse"]
        emb = self._embedding(models, model)
        n, S = rays.shape[0], self.N_samples
        depth_pass = depth is not None
        nerf = getattr(model, "variant", None) == "nerf"
        sc_lambda = getattr(self.cfgs.pipeline, "sc_lambda", 0.0)
        sc = sc_lambda > 0 and not depth_pass and not nerf
        self._calls += 1
        z, enc, enc_sc, aux, sky = encode_rays(model, emb, rays, extras, S, u=opts.get("u"), z=opts.get("z_vals"),
                                               seed=int(opts.get("seed", self._calls)),
                                               ray_offset=int(opts.get("ray_offset", 0)), want_sc=sc)
        C = model.semantic_n_classes
        terms = torch.zeros(8, dtype=torch.float32, device=rays.device)
        inv_n = 1.0 / max(int(global_rays) if global_rays is not None else n, 1)
        flags = (COMPOSITE_NO_CLAMP if nerf else 0) | (COMPOSITE_BETA_S if getattr(model, "beta_s", 0) else 0)
        p = _lib.LossParams(mode=2 if depth_pass else 0, color=1 if color == "satnerf" else 0, beta_min=beta_min,
                            inv_n=inv_n, lambda_s=lambda_s, ignore_index=ignore_index, lambda_c=lambda_c,
                            car_label=car_label, lambda_sc=sc_lambda, lambda_ds=lambda_ds, flags=flags,
                            sem_unc=int(semantic_uncertainty) if (semantic is not None and not depth_pass) else 0)
        counts = mask = None
        if semantic is not None and not depth_pass:
            semantic = as_labels(semantic)          # the reference's dataset yields uint8 labels; the kernel reads int64
            mask = as_ray_mask(ignore_mask)
            counts = label_counts(semantic, mask, C, ignore_index, car_label)   # same predicates as the loss kernel
            if reduce_counts is not None:
                reduce_counts(counts)
        out = mlp_rays(model, emb, enc, aux, sky, extras, n, S, HEADS_DEPTH if depth_pass else HEADS_ALL).view(n, S, -1)
        loss = CompositeLoss.apply(out, z, 0 if depth_pass else C, p, rgbs, None if depth_pass else semantic, depth,
                                   depth_weights, counts, terms, mask, reduce_counts)
        if sc:
            p_sc = _lib.LossParams(mode=1, color=0, beta_min=beta_min, inv_n=inv_n, lambda_s=0.0, ignore_index=-100,
                                   lambda_c=0.0, car_label=-1, lambda_sc=sc_lambda, lambda_ds=0.0, flags=0, sem_unc=0)
            out_sc = mlp_rays(model, emb, enc_sc, aux, None, extras, n, S, HEADS_SOLAR).view(n, S, -1)
            loss = loss + CompositeLoss.apply(out_sc, z, 0, p_sc, None, None, None, None, None, terms)
        self.last_counts = counts
        return loss, terms

    @staticmethod
    def _embedding(models, model):
        """the per-image embedding table(s) as the kernels take them: models["t"].weight, or - `use_separate_tj_for_semantic`
        (semantic/components/rendering.py:42-45) - [t | t_s] side by side as one (vocab, 2 tau) table (a differentiable cat:
        the gradient of the combined table splits back into the two embeddings)"""
        if "t" not in models:
            return None
        if getattr(model, "sep_ts", False):
            if "t_s" not in models:
                raise _lib.SnbError('use_separate_tj_for_semantic: models["t_s"] is missing (semantic/pipelines/rs_semantic.py:72-77)')
            return torch.cat([models["t"].weight, models["t_s"].weight], 1)
        return models["t"].weight

    @staticmethod
    def _empty_result(model, rays, S, sc):
        f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=rays.device)
        res = {"rgb": f(0, 3), "depth": f(0), "weights": f(0, S), "transparency": f(0, S), "albedo": f(0, S, 3),
               "sun": f(0, S, 1), "sky": f(0, S, 3), "beta": f(0, S, 1), "sigmas": f(0, S)}
        if model.kind == MODEL_SEMANTIC:
            res["semantic_logits"] = f(0, model.semantic_n_classes)
            res["semantic_label"] = torch.empty(0, dtype=torch.int64, device=rays.device)
            if getattr(model, "beta_s", 0):
                res["beta_semantic"] = f(0, S, 1)
        if getattr(model, "variant", None) == "snerf":
            del res["beta"], res["sigmas"]
        if getattr(model, "variant", None) == "nerf":
            res = {k: res[k] for k in ("rgb", "depth", "weights", "transparency")}
        if sc:
            res.update(weights_sc=f(0, S), transparency_sc=f(0, S), sun_sc=f(0, S, 1))
        res["_z_vals"] = f(0, S)
        return res


# the reference's two renderer class names, for configs / code that instantiate them by name
class SatNeRFB200Rendering(B200Renderer):
    pass


class NeRFB200Rendering(B200Renderer):
    """≙ baseline.components.rendering.NeRFRendering (rendering.py:103-118): models = {"coarse": NeRFB200}."""


class SNeRFB200Rendering(B200Renderer):
    """≙ baseline.components.rendering.SNeRFRendering (rendering.py:70-100): models = {"coarse": ShadowNeRFB200}."""


class RSSemanticB200Rendering(B200Renderer):
    """≙ semantic.components.rendering.RSSemanticRendering (rendering.py:14-16).  The reference's constructor takes a
    replacement `inference` callable; here sampling, the MLP and the compositing are fused CUDA kernels, so there is no
    per-stage hook to inject into - a non-None `inference` is refused instead of being silently ignored."""

    def __init__(self, cfgs, inference=None):
        if inference is not None:
            raise _lib.SnbError("RSSemanticB200Rendering does not accept a replacement `inference` callable: the B200 "
                                "path runs sampling, MLP and compositing as fused kernels (use the reference's "
                                "RSSemanticRendering for a custom inference function)")
        super().__init__(cfgs)

"""One optimisation step of the SatNeRF / Semantic-NeRF pipelines on the B200 path.

Mirrors the control flow of the reference's training steps
(baseline/components/training_step.py:19-59, semantic/components/training_step.py:10-99):
rgb batch -> render (main + solar-correction pass) -> colour loss (SNerfLoss before
``first_beta_epoch``, SatNerfLoss after) -> optional depth-supervision batch -> semantic CE ->
optional car regularisation -> backward -> Adam (lr 5e-4, base_ray_pipeline.py:246-269).
Lightning is not part of the path (and not installed here): this is the plain loop the benchmark
and the smoke test drive.  Data parallel: each rank renders its shard of the global batch; one
bucketed gradient all-reduce per step (dist.py).
"""
from __future__ import annotations

import types
from typing import Dict, Optional

import torch

from . import _lib, dist as snb_dist
from ._lib import check, ptr, stream
from .losses import DepthLoss, SatNerfLoss, SemanticCarRegLoss, SemanticLoss, SNerfLoss
from .model import NeRFB200, RSSemanticNeRFB200, SatNeRFB200, ShadowNeRFB200
from .renderer import B200Renderer


def default_cfgs(kind: str = "semantic", n_samples: int = 64, sc_lambda: float = 0.05, **over):
    """cfgs.pipeline with the field names and defaults of configs/pipelines/{satnerf,rs_semantic}.toml."""
    p = dict(n_samples=n_samples, render_chunk_size=40960, batch_size=1024, learnrate=5e-4, fc_units=512, fc_layers=8,
             fc_skips=[4], fc_use_full_features=False, activation_function="siren", mapping_pos_n_freq=10,
             mapping_dir_n_freq=4, sc_lambda=sc_lambda, depth_enabled=True, depth_supervision_drop=0.25, ds_lambda=1000,
             first_beta_epoch=2, t_embedding_vocab=50, t_embedding_tau=4, ds_noweights=False, lambda_s=0.04,
             semantic_activation_function="sigmoid", use_tj_for_s=False, use_tj_instead_of_beta=False,
             use_beta_for_s=False, detach_beta_for_s=False, use_separate_beta_for_s=False,
             use_separate_tj_for_semantic=False, ignore_car_index=True, use_car_reg_loss=False, car_reg_loss_start=3,
             lambda_c=0.1)
    p.update(over)
    return types.SimpleNamespace(pipeline=types.SimpleNamespace(**p))


class Trainer:
    def __init__(self, cfgs, kind: str = "semantic", n_classes: int = 6, device="cuda", car_index: int = 4,
                 world: int = 1, rank: int = 0, seed: int = 0, fused_loss: bool = True):
        p = cfgs.pipeline
        self.cfgs, self.kind, self.device, self.world, self.rank = cfgs, kind, torch.device(device), world, rank
        torch.manual_seed(seed)  # identical initial replicas on every rank
        if kind == "semantic":
            model = RSSemanticNeRFB200(cfgs, types.SimpleNamespace(semantic_n_classes=n_classes))
        elif kind == "nerf":    # baseline/pipelines/nerf.py:23-40: NerfLoss (MSE), NeRFTrainingStep, no embedding, no solar pass
            model = NeRFB200(layers=p.fc_layers, feat=p.fc_units, skips=p.fc_skips)
        elif kind == "snerf":   # baseline/pipelines/snerf.py:21-38: SNerfLoss, NeRFTrainingStep, no embedding, no depth batch
            model = ShadowNeRFB200(layers=p.fc_layers, feat=p.fc_units, skips=p.fc_skips)
        else:
            model = SatNeRFB200(cfgs, layers=p.fc_layers, feat=p.fc_units, skips=p.fc_skips,
                                t_embedding_dims=p.t_embedding_tau)
        self.models = {"coarse": model.to(self.device)}
        if kind not in ("snerf", "nerf"):
            self.models["t"] = torch.nn.Embedding(p.t_embedding_vocab, p.t_embedding_tau).to(self.device)
        self.renderer = B200Renderer(cfgs)
        self.loss = SatNerfLoss(lambda_sc=p.sc_lambda)
        self.loss_without_beta = SNerfLoss(lambda_sc=p.sc_lambda)
        self.depth_loss = DepthLoss(lambda_ds=p.ds_lambda)
        self.car_index = car_index
        if kind == "semantic":
            self.semantic_loss = SemanticLoss(p.lambda_s, car_index, ignore_car_index=p.ignore_car_index)
            self.car_reg_loss = SemanticCarRegLoss(p.lambda_c, car_index) if p.use_car_reg_loss else None
        self.lr, self.betas, self.eps = p.learnrate, (0.9, 0.999), 1e-8
        flat = model.flat
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.emb_opt = torch.optim.Adam(self.models["t"].parameters(), lr=self.lr) if "t" in self.models else None
        self.step_idx = 0
        # fused_loss: compositing + the loss modules + their backward in one kernel per pass (renderer.render_loss,
        # SURVEY 8f rank 1); False runs render_rays() + the reference-shaped loss modules (what a Lightning pipeline does)
        self.fused_loss = fused_loss
        self.reducer = snb_dist.GradAllReducer(snb_dist.bucket_ranges(model.table, flat.numel(), 3))

    # -- one step ---------------------------------------------------------------------------------------
    def training_step(self, batch: Dict[str, torch.Tensor], epoch: int = 2, depth_batch: Optional[dict] = None,
                      ray_offset: int = 0) -> torch.Tensor:
        p = self.cfgs.pipeline
        model, emb = self.models["coarse"], self.models.get("t")
        self.step_idx += 1
        opts = {"seed": self.step_idx, "ray_offset": ray_offset}
        if self.fused_loss:
            return self._fused_step(batch, epoch, depth_batch, opts)
        results = self.renderer.render_rays(self.models, batch["rays"], batch["extras"], epoch=epoch, render_options=opts)
        if epoch < p.first_beta_epoch or self.kind in ("snerf", "nerf"):
            loss, loss_dict = self.loss_without_beta(results, batch["rgbs"])
        else:
            loss, loss_dict = self.loss(results, batch["rgbs"])
        if depth_batch is not None:
            tmp = self.renderer.render_rays(self.models, depth_batch["rays"], depth_batch["extras"], epoch=epoch,
                                            render_options={"seed": self.step_idx + (1 << 20), "heads": "depth"})
            w = 1.0 if p.ds_noweights else depth_batch["weights"].flatten()
            l_d, d = self.depth_loss(tmp, depth_batch["depths"].flatten(), w)
            loss = loss + l_d
            loss_dict.update(d)
        if self.kind == "semantic":
            l_s, d = self.semantic_loss(results, batch["semantic"])
            loss = loss + l_s
            loss_dict.update(d)
            if self.car_reg_loss is not None and epoch >= p.car_reg_loss_start:
                l_c, d = self.car_reg_loss(results, batch["semantic"])
                loss = loss + l_c
                loss_dict.update(d)
        model.flat.grad = None
        if emb is not None:
            emb.weight.grad = None
        loss.backward()
        self.optimizer_step()
        self.last_loss_dict = loss_dict
        return loss.detach()

    def _fused_step(self, batch, epoch, depth_batch, opts):
        """the same step through renderer.render_loss: same loss value and gradients, no per-sample output tensors"""
        p = self.cfgs.pipeline
        model, emb = self.models["coarse"], self.models.get("t")
        sem = self.kind == "semantic"
        car = sem and self.car_reg_loss is not None and epoch >= p.car_reg_loss_start
        loss, terms = self.renderer.render_loss(
            self.models, batch["rays"], batch["extras"], batch["rgbs"], batch["semantic"] if sem else None,
            color="snerf" if (epoch < p.first_beta_epoch or self.kind in ("snerf", "nerf")) else "satnerf",
            lambda_s=p.lambda_s if sem else 0.0, ignore_index=self.car_index if (sem and p.ignore_car_index) else -100,
            lambda_c=p.lambda_c if car else 0.0, car_label=self.car_index, render_options=opts)
        if depth_batch is not None:
            w = None if p.ds_noweights else depth_batch["weights"].flatten()
            l_d, t_d = self.renderer.render_loss(self.models, depth_batch["rays"], depth_batch["extras"], None,
                                                 depth=depth_batch["depths"].flatten(), depth_weights=w, lambda_ds=p.ds_lambda,
                                                 render_options={"seed": self.step_idx + (1 << 20)})
            loss, terms = loss + l_d, terms + t_d
        model.flat.grad = None
        if emb is not None:
            emb.weight.grad = None
        loss.backward()
        self.optimizer_step()
        self.last_loss_terms = terms          # device tensor, order of autograd.LOSS_TERMS (no host sync here)
        return loss.detach()

    def optimizer_step(self):
        model, emb = self.models["coarse"], self.models.get("t")
        g = model.flat.grad
        if self.world > 1:
            self.reducer.launch(g)
            if emb is not None:
                torch.distributed.all_reduce(emb.weight.grad)
                emb.weight.grad.mul_(1.0 / self.world)
            self.reducer.wait()
        lib = _lib.load()
        check(lib.snb_adam_step(ptr(model.flat.data), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq), g.numel(),
                                self.lr, self.betas[0], self.betas[1], self.eps, self.step_idx, 1.0 / self.world,
                                stream()), "snb_adam_step")
        model.mark_dirty()
        if self.emb_opt is not None:
            self.emb_opt.step()

    # -- chunked no-grad render of a whole image (BaseRayPipeline.forward / batched_inference) ---------------
    @torch.no_grad()
    def render_image(self, rays, extras, chunk: Optional[int] = None, keys=("rgb_coarse", "depth_coarse",
                                                                           "semantic_label_coarse"),
                     heads: str = "all"):
        """eval/utils/util.py:13-42 with preallocated outputs instead of the quadratic torch.cat."""
        chunk = chunk or self.cfgs.pipeline.render_chunk_size
        n = rays.shape[0]
        out: Dict[str, torch.Tensor] = {}
        want_sc = any("_sc_" in k for k in keys)   # the solar-correction pass only produces the *_sc keys
        for i in range(0, n, chunk):
            res = self.renderer.render_rays(self.models, rays[i:i + chunk], extras[i:i + chunk],
                                            render_options={"seed": 1, "ray_offset": i, "heads": heads,
                                                            "solar_pass": want_sc})
            for k in keys:
                if k not in res:
                    continue
                if k not in out:
                    out[k] = torch.empty((n,) + tuple(res[k].shape[1:]), dtype=res[k].dtype, device=res[k].device)
                out[k][i:i + chunk] = res[k]
        return out

"""One optimisation step of the SatNeRF / Semantic-NeRF pipelines on the B200 path.

Mirrors the control flow of the reference's training steps
(baseline/components/training_step.py:19-59, semantic/components/training_step.py:10-99):
rgb batch -> render (main + solar-correction pass) -> colour loss (SNerfLoss before
``first_beta_epoch``, SatNerfLoss after) -> optional depth-supervision batch -> semantic CE (with the
dataset's ``semantic_sparsity_mask``) -> optional car regularisation -> backward -> Adam (lr 5e-4,
base_ray_pipeline.py:246-269).  Lightning is not part of the path (and not installed here): this is the
plain loop the benchmark and the smoke test drive.

Three ways to run the same step (same loss, same gradients - tests/test_gpu_step.py):
  fused_loss=False            render_rays() + the reference-shaped loss modules under autograd (what a
                              Lightning pipeline does with the plug-in renderer);
  fused_loss=True, direct=False   renderer.render_loss() under autograd (K3 + losses fused);
  fused_loss=True, direct=True    (default) the kernels of the step called back to back on persistent
                              buffers - no autograd graph, no per-step allocation, one flat gradient
                              buffer - optionally replayed as ONE CUDA graph (graph=True, single GPU).

Data parallel: each rank renders its shard of the global batch; the loss means run over the
GLOBAL batch (global ray count, all-reduced masked-mean denominators), so the per-rank gradients SUM to
the single-process gradient of the concatenated batch; the flat gradient is all-reduced in three buckets
(heads, late trunk, early trunk + embedding), each started on NCCL's stream as soon as the backward
pass has finished it (events recorded inside snb_mlp_backward) while the remaining weight-gradient
GEMMs still run.
"""
from __future__ import annotations

import ctypes as C
import types
from typing import Dict, Optional

import torch

from . import _lib, dist as snb_dist
from ._lib import COMPOSITE_BETA_S, COMPOSITE_NO_CLAMP, HEADS_ALL, HEADS_DEPTH, HEADS_SOLAR, check, ptr, stream
from .autograd import as_labels, as_ray_mask, t_steps
from .losses import (DepthLoss, NerfLoss, SatNerfLoss, SemanticCarRegLoss, SemanticLoss, SemanticUncertaintyLoss,
                     SNerfLoss)
from .model import NeRFB200, RSSemanticNeRFB200, SatNeRFB200, ShadowNeRFB200
from .renderer import B200Renderer

EMB_PAD = 2048  # floats reserved in front of the model parameters for the embedding tables: t at 0, t_s (if any) at 1024
                # (vocabulary 50 x t_embedding_tau <= 12 = 600 floats per table)


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


class _PassBuffers:
    """Persistent device buffers of one ray batch size for the direct step (K1 outputs, the MLP training workspace, packed
    head outputs and their gradients).  ~25 KB per sample and pass.  With a solar-correction pass its points are rows
    [P, 2P) of the SAME enc / workspace / out / g_out buffers (snb_mlp_forward_with_solar): the weight gradients of the layers
    both passes share then run once over all 2P rows."""

    def __init__(self, model, n: int, S: int, dev, want_sc: bool, has_emb: bool, depth_only: bool):
        lib = _lib.load()
        P = n * S
        f32 = dict(dtype=torch.float32, device=dev)
        bf16 = dict(dtype=torch.bfloat16, device=dev)
        nerf = model.kind == _lib.MODEL_NERF
        self.n, self.P = n, P
        self.rays = torch.empty(n, 8, **f32)
        self.extras = torch.empty(n, 4, **f32)
        self.z = torch.empty(n, S, **f32)
        rows = 2 * P if want_sc else P
        self.enc_all = torch.empty(rows, model.enc_ld, **bf16)
        self.enc = self.enc_all[:P]
        self.enc_sc = self.enc_all[P:] if want_sc else None
        self.aux = torch.empty(P, 16, **bf16)
        self.aux32 = torch.empty(P, 32, **bf16) if nerf else None
        self.sky = None if (nerf or depth_only) else torch.empty(n, 3, **f32)
        nbytes = lib.snb_mlp_workspace_bytes(model._h, rows, 1)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        n_out = model.n_out_kernel
        self.out_all = torch.empty(rows, n_out, **f32)
        self.g_out_all = torch.empty(rows, n_out, **f32)
        self.out, self.g_out = self.out_all[:P], self.g_out_all[:P]
        self.out_sc = self.out_all[P:] if want_sc else None
        self.g_out_sc = self.g_out_all[P:] if want_sc else None
        self.g_aux = torch.empty(P, 16, **f32) if (has_emb and not depth_only) else None
        # targets (static addresses, so the step can be replayed as a CUDA graph)
        self.rgbs = torch.empty(n, 3, **f32)
        self.labels = torch.empty(n, dtype=torch.int64, device=dev)
        self.mask = torch.empty(n, dtype=torch.uint8, device=dev)
        self.depths = torch.empty(n, **f32)
        self.dweights = torch.empty(n, **f32)


class Trainer:
    def __init__(self, cfgs, kind: str = "semantic", n_classes: int = 6, device="cuda", car_index: int = 4,
                 world: int = 1, rank: int = 0, seed: int = 0, fused_loss: bool = True, direct: bool = True,
                 graph: bool = False, micro_batch: Optional[int] = None):
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
        if kind == "semantic" and getattr(p, "use_separate_tj_for_semantic", False):
            self.models["t_s"] = torch.nn.Embedding(p.t_embedding_vocab, p.t_embedding_tau).to(self.device)
        emb = self.models.get("t")
        # ONE flat buffer [embedding table (padded) | model parameters] for parameters, gradients and Adam moments: the
        # optimiser is one launch, and the embedding gradient rides the last all-reduce bucket instead of its own collective
        self.n_emb = emb.weight.numel() if emb is not None else 0
        if self.n_emb > EMB_PAD // 2:
            raise _lib.SnbError(f"embedding table of {self.n_emb} floats exceeds the {EMB_PAD // 2} reserved")
        n_params = model.flat.numel()
        self.pbuf = torch.zeros(EMB_PAD + n_params, dtype=torch.float32, device=self.device)
        self.pbuf[EMB_PAD:].copy_(model.flat.data)
        model.flat.data = self.pbuf[EMB_PAD:]
        if emb is not None:
            self.pbuf[:self.n_emb].copy_(emb.weight.data.reshape(-1))
            emb.weight.data = self.pbuf[:self.n_emb].view_as(emb.weight)
        emb_s = self.models.get("t_s")
        if emb_s is not None:
            self.pbuf[EMB_PAD // 2:EMB_PAD // 2 + self.n_emb].copy_(emb_s.weight.data.reshape(-1))
            emb_s.weight.data = self.pbuf[EMB_PAD // 2:EMB_PAD // 2 + self.n_emb].view_as(emb_s.weight)
        self.gbuf = torch.zeros_like(self.pbuf)
        self.exp_avg = torch.zeros_like(self.pbuf)
        self.exp_avg_sq = torch.zeros_like(self.pbuf)
        self.renderer = B200Renderer(cfgs)
        nerf = kind == "nerf"
        sc_lambda = 0.0 if nerf else getattr(p, "sc_lambda", 0.0)
        self.loss = SatNerfLoss(lambda_sc=sc_lambda)
        # NeRF: NerfLoss = plain MSE, no solar term (baseline/pipelines/nerf.py:23-24; nerf.toml has no sc_lambda)
        self.loss_without_beta = NerfLoss() if nerf else SNerfLoss(lambda_sc=sc_lambda)
        self.depth_loss = DepthLoss(lambda_ds=p.ds_lambda)
        self.car_index = car_index
        if kind == "semantic":
            self.semantic_loss = SemanticLoss(p.lambda_s, car_index, ignore_car_index=p.ignore_car_index)
            self.car_reg_loss = SemanticCarRegLoss(p.lambda_c, car_index) if p.use_car_reg_loss else None
            self.uncertainty_semantic_loss = SemanticUncertaintyLoss(p.lambda_s, car_index,
                                                                     detach_beta_for_s=getattr(p, "detach_beta_for_s", False),
                                                                     ignore_car_index=p.ignore_car_index)
        self.lr, self.betas, self.eps = p.learnrate, (0.9, 0.999), 1e-8
        self.step_idx = 0
        # fused_loss: compositing + the loss modules + their backward in one kernel per pass; False runs
        # render_rays() + the reference-shaped loss modules (what a Lightning pipeline does)
        self.fused_loss = fused_loss
        self.direct = direct and fused_loss
        if emb_s is not None:
            # use_separate_tj_for_semantic: K1 takes the two tables side by side as one (vocab, 2 tau) table (renderer.py
            # _embedding); the direct step keeps that table and its gradient in two small persistent buffers
            vocab, tau = emb.weight.shape
            self._ew2 = torch.empty(vocab, 2 * tau, dtype=torch.float32, device=self.device)
            self._g_ew2 = torch.zeros(vocab, 2 * tau, dtype=torch.float32, device=self.device)
        self.use_graph = bool(graph) and self.direct and world == 1
        # micro_batch: the direct step runs batches larger than this many rays as several forward / backward passes that
        # accumulate into the one gradient buffer before the single optimiser step (the saved activations cost ~25 KB per
        # sample and pass: 8192 rays x 64 samples = 26 GB; a 65 536-ray global batch does not fit one GPU in one piece)
        self.micro_batch = micro_batch
        # gradient buckets in completion order (snb_model_grad_buckets), as ranges of gbuf; the last one also carries the
        # embedding gradient
        lo, hi = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        check(_lib.load().snb_model_grad_buckets(model._h, lo, hi), "snb_model_grad_buckets")
        self.buckets = [(EMB_PAD + lo[b], EMB_PAD + hi[b]) for b in range(3)]
        self.buckets[2] = (0, self.buckets[2][1])
        self._bufs: Dict[tuple, _PassBuffers] = {}
        self._graphs: Dict[tuple, tuple] = {}
        self._events = None
        self._side = None
        self._counts = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._terms = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._seed_dev = torch.zeros(2, dtype=torch.int64, device=self.device)    # [rgb batch key, depth batch key]
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.last_loss_terms = self._terms
        self.last_loss_dict = {}

    # -- gradient views for code that reads .grad (tests, a user's own optimiser) -----------------------------
    def _bind_grads(self):
        model, emb = self.models["coarse"], self.models.get("t")
        model.flat.grad = self.gbuf[EMB_PAD:]
        if emb is not None:
            emb.weight.grad = self.gbuf[:self.n_emb].view_as(emb.weight)
        if "t_s" in self.models:
            w = self.models["t_s"].weight
            w.grad = self.gbuf[EMB_PAD // 2:EMB_PAD // 2 + self.n_emb].view_as(w)

    # -- one step ---------------------------------------------------------------------------------------
    def training_step(self, batch: Dict[str, torch.Tensor], epoch: int = 2, depth_batch: Optional[dict] = None,
                      ray_offset: Optional[int] = None, global_rays: Optional[int] = None,
                      global_depth_rays: Optional[int] = None, depth_ray_offset: Optional[int] = None) -> torch.Tensor:
        """`batch`: rays (n,8), extras (n,4), rgbs (n,3) [, semantic (n,) or (n,1) any integer dtype,
        semantic_sparsity_mask (n,) bool].  Data parallel: `batch` is this rank's shard, `global_rays` the size of the
        global batch (default world * n) and `ray_offset` the shard's first ray index in it (the Philox key of a ray is its
        index in the global batch; default rank * n); likewise `global_depth_rays` / `depth_ray_offset` for the depth batch."""
        self.step_idx += 1
        n = batch["rays"].shape[0]
        if ray_offset is None:
            ray_offset = batch.get("_ray_offset", self.rank * n)
        if global_rays is None:
            global_rays = batch.get("_global_rays", self.world * n)
        if depth_batch is not None:
            nd = depth_batch["rays"].shape[0]
            if global_depth_rays is None:
                global_depth_rays = depth_batch.get("_global_rays", self.world * nd)
            if depth_ray_offset is None:
                depth_ray_offset = depth_batch.get("_ray_offset", self.rank * nd)
        offs = (int(ray_offset), int(depth_ray_offset or 0))
        if self.direct:
            return self._direct_step(batch, epoch, depth_batch, offs, int(global_rays), global_depth_rays)
        return self._autograd_step(batch, epoch, depth_batch, offs, int(global_rays), global_depth_rays)

    @staticmethod
    def _depths(depth_batch):
        dd = depth_batch["depths"]            # semantic/components/training_step.py:39: torch.flatten(depths[:, 0])
        return (dd[:, 0] if dd.dim() == 2 else dd).flatten()

    def _loss_config(self, epoch):
        p = self.cfgs.pipeline
        sem = self.kind == "semantic"
        car = sem and self.car_reg_loss is not None and epoch >= p.car_reg_loss_start
        color = "snerf" if (epoch < self._first_beta_epoch() or self.kind in ("snerf", "nerf")) else "satnerf"
        return sem, car, color

    def _first_beta_epoch(self):
        """`use_tj_instead_of_beta` disables the beta loss for good (semantic/pipelines/rs_semantic.py:28-33)"""
        p = self.cfgs.pipeline
        return 10 ** 7 if getattr(p, "use_tj_instead_of_beta", False) else p.first_beta_epoch

    def _sem_unc(self, epoch) -> int:
        """which semantic loss the step uses (semantic/components/training_step.py:50-75): the plain cross-entropy before
        `first_beta_epoch` or without `use_beta_for_s`, else the uncertainty-weighted one (2: beta detached)"""
        p = self.cfgs.pipeline
        if self.kind != "semantic" or epoch < self._first_beta_epoch() or not getattr(p, "use_beta_for_s", False):
            return 0
        return 2 if getattr(p, "detach_beta_for_s", False) else 1

    def _reduce_counts(self, counts):
        if self.world > 1:
            torch.distributed.all_reduce(counts)

    def _autograd_step(self, batch, epoch, depth_batch, offs, global_rays, global_depth_rays):
        p = self.cfgs.pipeline
        sem, car, color = self._loss_config(epoch)
        opts = {"seed": self.step_idx, "ray_offset": offs[0]}
        mask = batch.get("semantic_sparsity_mask")
        self.gbuf.zero_()
        self._bind_grads()
        if self.fused_loss:
            loss, terms = self.renderer.render_loss(
                self.models, batch["rays"], batch["extras"], batch["rgbs"], batch["semantic"] if sem else None,
                color=color, lambda_s=p.lambda_s if sem else 0.0,
                ignore_index=self.car_index if (sem and p.ignore_car_index) else -100,
                lambda_c=p.lambda_c if car else 0.0, car_label=self.car_index, ignore_mask=mask, global_rays=global_rays,
                reduce_counts=self._reduce_counts, semantic_uncertainty=self._sem_unc(epoch), render_options=opts)
            if depth_batch is not None:
                w = None if p.ds_noweights else depth_batch["weights"].flatten()
                l_d, t_d = self.renderer.render_loss(self.models, depth_batch["rays"], depth_batch["extras"], None,
                                                     depth=self._depths(depth_batch), depth_weights=w,
                                                     lambda_ds=p.ds_lambda, global_rays=global_depth_rays,
                                                     render_options={"seed": self.step_idx + (1 << 20),
                                                                     "ray_offset": offs[1]})
                loss, terms = loss + l_d, terms + t_d
            self.last_loss_terms = terms          # device tensor, order of autograd.LOSS_TERMS (no host sync here)
        else:
            if self.world > 1:
                raise _lib.SnbError("the module-loss path takes per-rank means; data-parallel training uses fused_loss=True "
                                    "(global-batch means)")
            results = self.renderer.render_rays(self.models, batch["rays"], batch["extras"], epoch=epoch, render_options=opts)
            loss, loss_dict = (self.loss_without_beta if color == "snerf" else self.loss)(results, batch["rgbs"])
            if depth_batch is not None:
                tmp = self.renderer.render_rays(self.models, depth_batch["rays"], depth_batch["extras"], epoch=epoch,
                                                render_options={"seed": self.step_idx + (1 << 20), "heads": "depth",
                                                                "ray_offset": offs[1]})
                w = 1.0 if p.ds_noweights else depth_batch["weights"].flatten()
                l_d, d = self.depth_loss(tmp, self._depths(depth_batch), w)
                loss = loss + l_d
                loss_dict.update(d)
            if sem:
                sl = self.uncertainty_semantic_loss if self._sem_unc(epoch) else self.semantic_loss
                l_s, d = sl(results, batch["semantic"], mask)
                loss = loss + l_s
                loss_dict.update(d)
                if car:
                    l_c, d = self.car_reg_loss(results, batch["semantic"], mask)
                    loss = loss + l_c
                    loss_dict.update(d)
            self.last_loss_dict = loss_dict
        loss.backward()    # accumulates in place into the bound views of gbuf
        self._all_reduce_and_step(None)
        return loss.detach()

    # -- the direct step: the kernels of the step back to back on persistent buffers ----------------------------
    def _buffers(self, tag: str, n: int, depth_only: bool) -> _PassBuffers:
        key = (tag, n)
        if key not in self._bufs:
            # two live batch sizes per role (the full micro-batch and a short last one); a third size replaces the oldest
            old = [k for k in self._bufs if k[0] == tag]
            if len(old) >= 2:
                del self._bufs[old[0]]
                self._graphs.clear()
            model = self.models["coarse"]
            sc = (not depth_only) and self.kind != "nerf" and getattr(self.cfgs.pipeline, "sc_lambda", 0.0) > 0
            self._bufs[key] = _PassBuffers(model, n, self.cfgs.pipeline.n_samples, self.device, sc, "t" in self.models,
                                           depth_only)
        return self._bufs[key]

    @staticmethod
    def _stage(dst: torch.Tensor, src: torch.Tensor):
        dst.copy_(src.reshape(dst.shape), non_blocking=True)

    def _direct_step(self, batch, epoch, depth_batch, ray_offset, global_rays, global_depth_rays):
        # ray_offset: (rgb batch, depth batch)
        p = self.cfgs.pipeline
        sem, car, color = self._loss_config(epoch)
        n = batch["rays"].shape[0]
        micro = self.micro_batch if (self.micro_batch and n > self.micro_batch) else n
        pieces = [(lo, min(lo + micro, n)) for lo in range(0, n, micro)]
        m = batch.get("semantic_sparsity_mask") if sem else None
        has_mask = m is not None
        d = None
        if depth_batch is not None:
            d = self._buffers("depth", depth_batch["rays"].shape[0], True)
            self._stage(d.rays, depth_batch["rays"])
            self._stage(d.extras, depth_batch["extras"])
            dd = depth_batch["depths"]
            self._stage(d.depths, dd[:, 0] if dd.dim() == 2 else dd)     # training_step.py:39: depths[:, 0]
            if not p.ds_noweights:
                self._stage(d.dweights, depth_batch["weights"])
        self._seed_dev[0].fill_(self.step_idx)
        self._seed_dev[1].fill_(self.step_idx + (1 << 20))
        self._step_dev.fill_(self.step_idx)
        counts_done = False
        sem_unc = self._sem_unc(epoch)
        if sem_unc and len(pieces) > 1:
            raise _lib.SnbError("use_beta_for_s needs the whole batch's statistics before any gradient: not available with "
                                "micro-batching (raise micro_batch or lower the batch size)")
        if len(pieces) > 1 and sem:
            # the masked-mean denominators run over the WHOLE batch (and, data parallel, over every rank's)
            self._counts.zero_()
            lab_all = as_labels(batch["semantic"].to(self.device, non_blocking=True))
            mask_all = as_ray_mask(m.to(self.device, non_blocking=True)) if has_mask else None
            check(_lib.load().snb_label_counts(ptr(lab_all), ptr(mask_all), n,
                                               self.models["coarse"].semantic_n_classes,
                                               self.car_index if p.ignore_car_index else -100, self.car_index,
                                               ptr(self._counts), stream()), "snb_label_counts")
            self._reduce_counts(self._counts)
            counts_done = True
        for i, (lo, hi) in enumerate(pieces):
            b = self._buffers("rgb", hi - lo, False)
            self._stage(b.rays, batch["rays"][lo:hi])
            self._stage(b.extras, batch["extras"][lo:hi])
            self._stage(b.rgbs, batch["rgbs"][lo:hi])
            if sem:
                self._stage(b.labels, batch["semantic"][lo:hi])   # any integer dtype (the dataset's uint8) -> int64
                if has_mask:
                    self._stage(b.mask, m[lo:hi])
            first, last = i == 0, i == len(pieces) - 1
            offs = (ray_offset[0] + lo, ray_offset[1])
            args = (b, d if first else None, sem, car, color, has_mask, offs, global_rays, global_depth_rays, first, last,
                    counts_done, n / max(global_rays, 1), sem_unc)
            cfg = (hi - lo, d.n if d is not None else 0, color, car, has_mask, offs, global_rays, global_depth_rays, sem_unc)
            if self.use_graph and len(pieces) == 1:
                entry = self._graphs.get(cfg)
                if entry is None:
                    # first call of a configuration runs eagerly (lazy one-time CUDA attribute calls, allocator warm-up) ...
                    self._graphs[cfg] = ("warm",)
                    self._run_direct(*args)
                elif entry[0] == "warm":
                    # ... the second is captured (and the capture replayed, since capturing does not execute)
                    lib = _lib.load()
                    g = torch.cuda.CUDAGraph()
                    l0 = lib.snb_profile_launch_count()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        self._run_direct(*args)
                    captured = lib.snb_profile_launch_count() - l0
                    lib.snb_profile_add_launches(-captured)            # captured, not yet executed
                    self._graphs[cfg] = entry = ("graph", g, captured, self._loss_out)
                if entry is not None and entry[0] == "graph":
                    entry[1].replay()
                    _lib.load().snb_profile_add_launches(entry[2])
                    self._loss_out = entry[3]
            else:
                self._run_direct(*args)
        self._bind_grads()
        self.last_loss_terms = self._terms
        return self._loss_out

    def _run_direct(self, b, d, sem, car, color, has_mask, ray_offset, global_rays, global_depth_rays, first=True, last=True,
                    counts_done=False, share=1.0, sem_unc=0):
        """K1 -> MLP forward (main, solar) -> K3 + losses (gradients of the packed rows) -> MLP backward (solar, main) ->
        ray-parameter gradients -> [depth batch the same way] -> all-reduce -> Adam -> re-pack.  Every buffer is persistent.
        first / last: this is the first / last micro-batch of the step (zero the accumulators / reduce and step);
        share: this rank's fraction of the global batch (its share of the constant of the log-beta term)."""
        lib = _lib.load()
        p = self.cfgs.pipeline
        model, emb = self.models["coarse"], self.models.get("t")
        S = p.n_samples
        st = stream()
        nerf = model.kind == _lib.MODEL_NERF
        sc_lambda = 0.0 if nerf else getattr(p, "sc_lambda", 0.0)
        Cn = model.semantic_n_classes
        n_out = model.n_out_kernel
        ew = emb.weight.detach() if emb is not None else None
        emb_s = self.models.get("t_s")
        if emb_s is not None:
            if first:
                torch.cat([ew, emb_s.weight.detach()], 1, out=self._ew2)
                self._g_ew2.zero_()
            ew = self._ew2
        vocab, tau = (ew.shape if ew is not None else (0, 0))
        if nerf:
            sw = (None, None, None, None)
            hidden = 0
        else:
            sw = model.sky_params()
            hidden = sw[0].shape[0]
        gflat = self.gbuf[EMB_PAD:]
        g_emb = self.gbuf[:self.n_emb] if emb is not None else None
        if emb_s is not None:
            g_emb = self._g_ew2
        ts = t_steps(S, self.device)
        if first:
            self._terms.zero_()
            self.gbuf.zero_()
        packed = model.packed()
        if self._events is None and self.world > 1:
            self._events = [torch.cuda.Event() for _ in range(3)]
            for e in self._events:
                e.record()                      # creates the underlying cudaEvent_t
            self._side = torch.cuda.Stream(device=self.device)
        ev_arr = None

        def encode(buf, seed_slot, want_sc, want_sky):
            check(lib.snb_sample_encode(ptr(buf.rays), ptr(buf.extras), None, 0, self._seed_dev[seed_slot:].data_ptr(),
                                        ray_offset[seed_slot],
                                        ptr(ts), ptr(ew), vocab, tau, ptr(sw[0]), ptr(sw[1]), ptr(sw[2]), ptr(sw[3]),
                                        hidden, buf.n, S, _lib.K1_KIND[model.kind], 0, ptr(buf.z), ptr(buf.enc),
                                        ptr(buf.enc_sc) if want_sc else None, ptr(buf.aux),
                                        ptr(buf.sky) if want_sky else None, st), "snb_sample_encode")
            if nerf:
                check(lib.snb_nerf_aux(buf.rays[:, 3:6].data_ptr(), 8, buf.n, S, ptr(buf.aux32), st), "snb_nerf_aux")

        def forward(buf, ws, enc, sky, mask, out):
            aux = buf.aux32 if nerf else buf.aux
            check(lib.snb_mlp_forward(model._h, ptr(packed), ptr(ws), ws.numel(), buf.P, ptr(enc), ptr(aux), ptr(sky), S, mask, 1,
                                      ptr(out), st), "snb_mlp_forward")

        def backward(buf, ws, enc, out, g_out, mask, g_aux, events):
            aux = buf.aux32 if nerf else buf.aux
            check(lib.snb_mlp_backward(model._h, ptr(packed), ptr(ws), ws.numel(), buf.P, ptr(enc), ptr(aux), ptr(out),
                                       ptr(g_out), mask, ptr(gflat), ptr(g_aux), events, st), "snb_mlp_backward")

        # ---- rgb batch ---------------------------------------------------------------------------------------
        sc = sc_lambda > 0
        encode(b, 0, sc, not nerf)
        counts = None
        work = None
        if sem:
            counts = self._counts
            if not counts_done:
                counts.zero_()
                check(lib.snb_label_counts(ptr(b.labels), ptr(b.mask) if has_mask else None, b.n, Cn,
                                           self.car_index if p.ignore_car_index else -100, self.car_index, ptr(counts), st),
                      "snb_label_counts")
                if self.world > 1:   # global masked-mean denominators; overlaps the forward passes
                    work = torch.distributed.all_reduce(counts, async_op=True)
        if sc:   # main rows [0, P) + solar rows [P, 2P) of one workspace
            check(lib.snb_mlp_forward_with_solar(model._h, ptr(packed), ptr(b.ws), b.ws.numel(), b.P, b.P, ptr(b.enc_all),
                                                 ptr(b.aux), ptr(b.sky), S, ptr(b.out_all), st), "snb_mlp_forward_with_solar")
        else:
            forward(b, b.ws, b.enc, b.sky, HEADS_ALL, b.out)
        if work is not None:
            work.wait()
        inv_n = 1.0 / max(global_rays, 1)
        lp = _lib.LossParams(mode=0, color=1 if color == "satnerf" else 0, beta_min=0.05, inv_n=inv_n,
                             lambda_s=p.lambda_s if sem else 0.0,
                             ignore_index=self.car_index if (sem and p.ignore_car_index) else -100,
                             lambda_c=p.lambda_c if car else 0.0, car_label=self.car_index, lambda_sc=sc_lambda,
                             lambda_ds=0.0, flags=(COMPOSITE_NO_CLAMP if nerf else 0) | (COMPOSITE_BETA_S if model.beta_s else 0),
                             sem_unc=sem_unc if sem else 0)
        if sem and sem_unc:
            # SemanticUncertaintyLoss = lambda_s * CE_mean * mean_r 1 / (2 beta_r^2): both batch means first (mode 3 pre-pass
            # into counts[4:6]; data parallel: summed over the ranks), then the main pass forms the gradients
            lp3 = _lib.LossParams.from_buffer_copy(lp)
            lp3.mode = 3
            check(lib.snb_composite_loss(ptr(b.out), ptr(b.z), b.n, S, n_out, Cn, None, ptr(b.labels),
                                         ptr(b.mask) if has_mask else None, None, None, ptr(counts), C.byref(lp3), None,
                                         counts[4:].data_ptr(), st), "snb_composite_loss (statistics pre-pass)")
            self._reduce_counts(counts[4:6])
        check(lib.snb_composite_loss(ptr(b.out), ptr(b.z), b.n, S, n_out, Cn, ptr(b.rgbs), ptr(b.labels) if sem else None,
                                     ptr(b.mask) if (sem and has_mask) else None, None, None, ptr(counts), C.byref(lp),
                                     ptr(b.g_out), ptr(self._terms), st), "snb_composite_loss")
        if sc:
            lp_sc = _lib.LossParams(mode=1, color=0, beta_min=0.05, inv_n=inv_n, lambda_s=0.0, ignore_index=-100, lambda_c=0.0,
                                    car_label=-1, lambda_sc=sc_lambda, lambda_ds=0.0, flags=0)
            check(lib.snb_composite_loss(ptr(b.out_sc), ptr(b.z), b.n, S, n_out, 0, None, None, None, None, None, None,
                                         C.byref(lp_sc), ptr(b.g_out_sc), ptr(self._terms), st), "snb_composite_loss")
        # ---- depth-supervision batch (semantic/components/training_step.py:31-49): trunk + sigma only ---------------
        if d is not None:
            encode(d, 1, False, False)
            forward(d, d.ws, d.enc, None, HEADS_DEPTH, d.out)
            lp_d = _lib.LossParams(mode=2, color=0, beta_min=0.05, inv_n=1.0 / max(int(global_depth_rays), 1), lambda_s=0.0,
                                   ignore_index=-100, lambda_c=0.0, car_label=-1, lambda_sc=0.0, lambda_ds=p.ds_lambda, flags=0)
            check(lib.snb_composite_loss(ptr(d.out), ptr(d.z), d.n, S, n_out, 0, None, None, None, ptr(d.depths),
                                         None if p.ds_noweights else ptr(d.dweights), None, C.byref(lp_d), ptr(d.g_out),
                                         ptr(self._terms), st), "snb_composite_loss")
            backward(d, d.ws, d.enc, d.out, d.g_out, HEADS_DEPTH, None, None)
        # ---- backward of the rgb batch: sky_color gradients first (they belong to the first bucket), solar pass, main pass
        if not nerf:
            check(lib.snb_ray_param_backward(model._h, ptr(model.flat.detach()), ptr(b.extras), ptr(b.sky), ptr(b.g_out), None,
                                             b.n, S, n_out, 0, 1, ptr(gflat), None, st), "snb_ray_param_backward")
        if self.world > 1 and last:
            ev_arr = (C.c_void_p * 3)(*[e.cuda_event for e in self._events])
        if sc:   # both passes' dgrad chains, then one weight-gradient GEMM per shared layer over all 2P rows
            check(lib.snb_mlp_backward_with_solar(model._h, ptr(packed), ptr(b.ws), b.ws.numel(), b.P, b.P, ptr(b.enc_all),
                                                  ptr(b.aux), ptr(b.out_all), ptr(b.g_out_all), ptr(gflat), ptr(b.g_aux),
                                                  ev_arr, st), "snb_mlp_backward_with_solar")
        else:
            backward(b, b.ws, b.enc, b.out, b.g_out, HEADS_ALL, b.g_aux, ev_arr)
        if b.g_aux is not None:   # embedding gradient: per-ray sums of the aux-column gradients, scattered by ts
            check(lib.snb_ray_param_backward(model._h, ptr(model.flat.detach()), ptr(b.extras), None, None, ptr(b.g_aux), b.n, S,
                                             n_out, tau, vocab, ptr(gflat), ptr(g_emb), st), "snb_ray_param_backward")
        if last and emb_s is not None:   # split the gradient of the side-by-side table back into the two embeddings
            h, t4 = EMB_PAD // 2, tau // 2
            self.gbuf[:self.n_emb].view(vocab, t4).add_(self._g_ew2[:, :t4])
            self.gbuf[h:h + self.n_emb].view(vocab, t4).add_(self._g_ew2[:, t4:])
        if last:
            # loss value: the terms' sum (+ the constant 3/2 of the log-beta term, baseline/components/loss.py:26); data
            # parallel, the per-rank values are shares that add up to the global-batch loss
            const = (1.5 * share if color == "satnerf" else 0.0) + \
                    (1.5 * p.lambda_s * share if (sem and sem_unc and model.beta_s) else 0.0)   # semantic log-beta_s (loss.py:27-30)
            self._loss_out = self._terms.sum() + const
            self._all_reduce_and_step(self._events if self.world > 1 else None, self._step_dev)

    # -- gradient all-reduce (sum: the losses are already normalised by the global batch) + Adam + re-pack -------------
    def _all_reduce_and_step(self, events, step_dev=None):
        """step_dev: device int holding the step number (the direct step keeps it current so that a captured CUDA graph
        replays with the right Adam bias corrections); None = use the host counter"""
        model = self.models["coarse"]
        lib = _lib.load()
        if self.world > 1:
            cur = torch.cuda.current_stream()
            works = []
            if events is not None:
                # buckets 0 and 1 start as soon as the backward pass has finished them; the last bucket (early trunk +
                # embedding) is final once everything launched so far has run
                last = torch.cuda.Event()
                last.record(cur)
                for bkt, ev in zip(self.buckets, (events[0], events[1], last)):
                    self._side.wait_event(ev)
                    with torch.cuda.stream(self._side):
                        works.append(torch.distributed.all_reduce(self.gbuf[bkt[0]:bkt[1]], async_op=True))
            else:
                for bkt in self.buckets:
                    works.append(torch.distributed.all_reduce(self.gbuf[bkt[0]:bkt[1]], async_op=True))
            for w in works:
                w.wait()                        # the current stream waits; the host does not
            if events is not None:
                cur.wait_stream(self._side)
        check(lib.snb_adam_step(ptr(self.pbuf), ptr(self.gbuf), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.pbuf.numel(),
                                self.lr, self.betas[0], self.betas[1], self.eps, self.step_idx, ptr(step_dev), 1.0,
                                stream()), "snb_adam_step")
        model.mark_dirty()
        if self.direct:
            model.packed()                      # re-pack inside the step (and inside its CUDA graph)

    # -- chunked no-grad render of a whole image (BaseRayPipeline.forward / batched_inference) ---------------
    @torch.no_grad()
    def render_image(self, rays, extras, chunk: Optional[int] = None, keys=("rgb_coarse", "depth_coarse",
                                                                           "semantic_label_coarse"),
                     heads: str = "all", seed: int = 1, ray_offset: int = 0, u=None):
        """eval/utils/util.py:13-42 (`batched_inference`) / baseline/pipelines/base_ray_pipeline.py:34-54 with preallocated
        outputs instead of the quadratic torch.cat.  The jitter is keyed on (seed, global ray index): the result does not
        depend on the chunk size or on how an image is split across ranks (`ray_offset` = first global ray of `rays`).
        `u` (N,S) replaces the jitter (parity tests)."""
        chunk = chunk or self.cfgs.pipeline.render_chunk_size
        n = rays.shape[0]
        out: Dict[str, torch.Tensor] = {}
        want_sc = any("_sc_" in k for k in keys)   # the solar-correction pass only produces the *_sc keys
        for i in range(0, n, chunk):
            opts = {"seed": seed, "ray_offset": ray_offset + i, "heads": heads, "solar_pass": want_sc}
            if u is not None:
                opts["u"] = u[i:i + chunk]
            res = self.renderer.render_rays(self.models, rays[i:i + chunk], extras[i:i + chunk], render_options=opts)
            for k in keys:
                if k not in res:
                    continue
                if k not in out:
                    out[k] = torch.empty((n,) + tuple(res[k].shape[1:]), dtype=res[k].dtype, device=res[k].device)
                out[k][i:i + chunk] = res[k]
        return out

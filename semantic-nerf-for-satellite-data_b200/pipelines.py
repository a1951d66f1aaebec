"""Pipeline plug-ins for a checkout of the reference (needs its own dependencies, e.g. Lightning).

The reference selects a pipeline by dotted class path in the pipeline TOML
(`pipeline = "semantic.pipelines.rs_semantic.RSSemanticPipeline"`, configs/pipelines/rs_semantic.toml:5;
resolved by framework/pipelines.py:341-352).  Pointing that line at
`semnerf_b200.pipelines.RSSemanticB200Pipeline` / `SatNeRFB200Pipeline` / `SNeRFB200Pipeline` swaps the B200 model and
renderer in; datasets, losses, training step, visualisers and checkpoints are the reference's own.
Only `_init_models` and `_init_renderer` are overridden (semantic/pipelines/rs_semantic.py:61-79,
baseline/pipelines/satnerf.py:51-69).

The classes are built lazily because the reference modules import pytorch_lightning, which is not
installed in the build image: `get_pipeline_classes()` raises ImportError with the missing module.
"""
from __future__ import annotations

import torch

from .model import NeRFB200, RSSemanticNeRFB200, SatNeRFB200, ShadowNeRFB200
from .renderer import NeRFB200Rendering, RSSemanticB200Rendering, SatNeRFB200Rendering, SNeRFB200Rendering

_CACHE = {}


def get_pipeline_classes():
    if _CACHE:
        return _CACHE
    from baseline.pipelines.satnerf import SatNeRFPipeline          # reference checkout on sys.path
    from baseline.pipelines.snerf import SNerfPipeline
    from semantic.pipelines.rs_semantic import RSSemanticPipeline

    from baseline.pipelines.nerf import NerfPipeline

    class NeRFB200Pipeline(NerfPipeline):
        def _init_models(self) -> dict:   # baseline/pipelines/nerf.py:26-34
            p = self.cfgs.pipeline
            return {"coarse": NeRFB200(layers=p.fc_layers, feat=p.fc_units, skips=p.fc_skips)}

        def _init_renderer(self):         # baseline/pipelines/nerf.py:36-37
            return NeRFB200Rendering(self.cfgs)

    class SNeRFB200Pipeline(SNerfPipeline):
        def _init_models(self) -> dict:   # baseline/pipelines/snerf.py:24-32
            p = self.cfgs.pipeline
            return {"coarse": ShadowNeRFB200(layers=p.fc_layers, feat=p.fc_units, skips=p.fc_skips)}

        def _init_renderer(self):         # baseline/pipelines/snerf.py:34-35
            return SNeRFB200Rendering(self.cfgs)

    class SatNeRFB200Pipeline(SatNeRFPipeline):
        def _init_models(self) -> dict:
            p = self.cfgs.pipeline
            return {"coarse": SatNeRFB200(self.cfgs, layers=p.fc_layers, feat=p.fc_units, skips=p.fc_skips,
                                          t_embedding_dims=p.t_embedding_tau),
                    "t": torch.nn.Embedding(p.t_embedding_vocab, p.t_embedding_tau)}

        def _init_renderer(self):
            return SatNeRFB200Rendering(self.cfgs)

    class RSSemanticB200Pipeline(RSSemanticPipeline):
        def _init_models(self) -> dict:
            p = self.cfgs.pipeline
            d = {"coarse": RSSemanticNeRFB200(self.cfgs, self.datasets["rgb"]),
                 "t": torch.nn.Embedding(p.t_embedding_vocab, p.t_embedding_tau)}
            if p.use_separate_tj_for_semantic:   # semantic/pipelines/rs_semantic.py:72-77
                d["t_s"] = torch.nn.Embedding(p.t_embedding_vocab, p.t_embedding_tau)
            return d

        def _init_renderer(self):
            return RSSemanticB200Rendering(self.cfgs)

    _CACHE.update(SatNeRFB200Pipeline=SatNeRFB200Pipeline, RSSemanticB200Pipeline=RSSemanticB200Pipeline,
                  SNeRFB200Pipeline=SNeRFB200Pipeline, NeRFB200Pipeline=NeRFB200Pipeline)
    return _CACHE


def __getattr__(name):  # `semnerf_b200.pipelines.RSSemanticB200Pipeline` resolves through importlib
    if name in ("SatNeRFB200Pipeline", "RSSemanticB200Pipeline", "SNeRFB200Pipeline", "NeRFB200Pipeline"):
        return get_pipeline_classes()[name]
    raise AttributeError(name)

"""The plug-in boundary exercised through the REFERENCE'S OWN pipeline machinery (CPU; skipped where /root/reference is
absent, e.g. on the GPU box): `framework.pipelines.load_pipeline` resolves `cfgs.pipeline.pipeline =
"semnerf_b200.pipelines.RSSemanticB200Pipeline"` (framework/pipelines.py:341-352), the class's `init_config` builds the
reference's pydantic config from configs/pipelines/rs_semantic.toml (framework/configs.py:71-75), and the reference's
`Pipeline.__init__` (framework/pipelines.py:22-46) drives `_init_datasets` / `_init_models` / `_init_renderer` /
`_init_loss` / `_init_training_step`.  Lightning, torchmetrics and the geo stack are absent from this image and are stubbed
(tests/ref_stubs.py); the datasets are replaced by a stand-in with the two attributes the pipelines read."""
import importlib
import os
import sys
import types

import pytest
import toml
import torch

REF = os.environ.get("SNB_REF", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "framework")), reason="needs the reference checkout")


class _Dataset:
    """what the pipelines read from their datasets at construction time: semantic_n_classes (rs_semantic.py:95), car_cls_idx
    (semantic/pipelines/rs_semantic.py:45-58), __len__ (base_ray_pipeline.py:256-260)"""
    semantic_n_classes = 6
    car_cls_idx = 4
    dataset_name = "stub"

    def __init__(self, cfgs, name, split):
        self.cfgs, self.name, self.split = cfgs, name, split

    def __len__(self):
        return 100000


@pytest.fixture(scope="module")
def ref():
    from tests import ref_stubs
    sys.path.insert(0, REF)
    ref_stubs.install()
    mods = types.SimpleNamespace(
        pipelines=importlib.import_module("framework.pipelines"),
        configs=importlib.import_module("framework.configs"),
        rs=importlib.import_module("semantic.pipelines.rs_semantic"),
        sat=importlib.import_module("baseline.pipelines.satnerf"),
        snerf=importlib.import_module("baseline.pipelines.snerf"),
        nerf=importlib.import_module("baseline.pipelines.nerf"))
    # the datasets read geo files at construction: stand-ins (the module-level names the pipelines instantiate)
    saved = []
    for mod, names in ((mods.rs, ("SemanticDataset", "SatNeRFDepthDataset")), (mods.sat, ("SatNeRFDataset", "SatNeRFDepthDataset")),
                       (mods.snerf, ("SNeRFDataset",)), (mods.nerf, ("NeRFDataset",))):
        for n in names:
            if hasattr(mod, n):
                saved.append((mod, n, getattr(mod, n)))
                setattr(mod, n, _Dataset)
    yield mods
    for mod, n, v in saved:
        setattr(mod, n, v)
    sys.path.remove(REF)
    for name in [n for n in sys.modules if n.split(".")[0] in ("framework", "baseline", "semantic", "eval", "data_prep")]:
        del sys.modules[name]
    import semnerf_b200.pipelines as P
    P._CACHE.clear()
    ref_stubs.uninstall()


def _cfgs(ref, toml_name, dotted):
    """MainConfig's recipe (framework/configs.py:62-75) without the run-directory sanity checks"""
    data = toml.load(os.path.join(REF, "configs", "pipelines", toml_name))
    data["pipeline"] = dotted
    name = dotted.split(".")
    cls = getattr(importlib.import_module(".".join(name[:-1])), name[-1])
    return types.SimpleNamespace(pipeline=cls.init_config(data), run=ref.configs.RunConfig(max_train_steps=1000))


@pytest.mark.parametrize("toml_name,ours,theirs,model_cls,renderer_cls", [
    ("rs_semantic.toml", "semnerf_b200.pipelines.RSSemanticB200Pipeline", "semantic.pipelines.rs_semantic.RSSemanticPipeline",
     "RSSemanticNeRFB200", "RSSemanticB200Rendering"),
    ("satnerf.toml", "semnerf_b200.pipelines.SatNeRFB200Pipeline", "baseline.pipelines.satnerf.SatNeRFPipeline",
     "SatNeRFB200", "SatNeRFB200Rendering"),
])
def test_reference_load_pipeline_builds_the_b200_plugin(ref, toml_name, ours, theirs, model_cls, renderer_cls):
    import semnerf_b200.model as M
    import semnerf_b200.renderer as R
    cfgs = _cfgs(ref, toml_name, ours)
    assert type(cfgs.pipeline).__name__ == type(_cfgs(ref, toml_name, theirs).pipeline).__name__   # the reference's own config class
    pipe = ref.pipelines.load_pipeline(cfgs)                     # framework/pipelines.py:341-352
    base = getattr(importlib.import_module(".".join(theirs.split(".")[:-1])), theirs.split(".")[-1])
    assert isinstance(pipe, base)                                # a subclass: everything but two factories is the reference's
    assert type(pipe.models["coarse"]) is getattr(M, model_cls) and type(pipe.renderer) is getattr(R, renderer_cls)
    assert isinstance(pipe.models["t"], torch.nn.Embedding) and tuple(pipe.models["t"].weight.shape) == (50, 4)
    assert pipe.model_coarse is pipe.models["coarse"]            # registered as a submodule (framework/pipelines.py:204-214)
    # the reference's losses and training step were set up by the inherited factories
    assert type(pipe.loss).__name__ == "SatNerfLoss" and type(pipe.depth_loss).__name__ == "DepthLoss"
    if "semantic" in toml_name:
        assert type(pipe.semantic_loss).__name__ == "SemanticLoss" and type(pipe._training_step).__name__ == "RSSemanticTrainingStep"
        assert pipe.models["coarse"].semantic_n_classes == 6
    # checkpoints interchange: the same state_dict keys and shapes as the reference pipeline built from the same TOML
    theirs_pipe = ref.pipelines.load_pipeline(_cfgs(ref, toml_name, theirs))
    sd, sd_ref = pipe.state_dict(), theirs_pipe.state_dict()
    assert list(sd.keys()) == list(sd_ref.keys())
    assert all(tuple(sd[k].shape) == tuple(sd_ref[k].shape) for k in sd)
    pipe.load_state_dict(sd_ref)                                 # a reference checkpoint loads into the plug-in ...
    assert torch.equal(pipe.state_dict()["model_coarse.fc_net.8.weight"], sd_ref["model_coarse.fc_net.8.weight"])
    theirs_pipe.load_state_dict(pipe.state_dict())               # ... and the plug-in's into the reference
    # the optimiser the reference configures sees the plug-in's parameters (base_ray_pipeline.py:246-269)
    opt = pipe.configure_optimizers()["optimizer"]
    n_opt = sum(p.numel() for g in opt.param_groups for p in g["params"])
    assert n_opt == sum(v.numel() for v in sd.values())
    # renderer signature the pipeline calls (base_ray_pipeline.py:34-52)
    import inspect
    want = list(inspect.signature(type(theirs_pipe.renderer).render_rays).parameters)
    assert list(inspect.signature(type(pipe.renderer).render_rays).parameters) == want


def test_plugin_renderer_refuses_an_injected_inference_function(ref):
    """RSSemanticRendering(cfgs, inference=fn) injects a replacement per-point function (semantic/components/rendering.py:
    14-16); the fused kernels have no such hook - a non-None argument raises instead of being dropped silently."""
    from semnerf_b200 import _lib
    from semnerf_b200.renderer import RSSemanticB200Rendering
    cfgs = _cfgs(ref, "rs_semantic.toml", "semnerf_b200.pipelines.RSSemanticB200Pipeline")
    RSSemanticB200Rendering(cfgs)
    with pytest.raises(_lib.SnbError):
        RSSemanticB200Rendering(cfgs, inference=lambda *a, **k: None)


def test_unsupported_configurations_raise_at_construction(ref):
    from semnerf_b200 import _lib
    for flag, value in (("t_embedding_tau", 13), ("fc_units", 256)):
        cfgs = _cfgs(ref, "rs_semantic.toml", "semnerf_b200.pipelines.RSSemanticB200Pipeline")
        setattr(cfgs.pipeline, flag, value)
        with pytest.raises(_lib.SnbError):
            ref.pipelines.load_pipeline(cfgs)


def test_full_features_and_embedding_width_build_the_same_state_dict_as_the_reference(ref):
    """fc_use_full_features = true, t_embedding_tau = 6, activation_function = "relu" and mapping_pos_n_freq = 6 together with
    every head variant: same state_dict keys and shapes as the reference pipeline built from the same config"""
    pipes = []
    for dotted in ("semnerf_b200.pipelines.RSSemanticB200Pipeline", "semantic.pipelines.rs_semantic.RSSemanticPipeline"):
        cfgs = _cfgs(ref, "rs_semantic.toml", dotted)
        for flag in ("use_tj_for_s", "use_separate_beta_for_s", "use_separate_tj_for_semantic", "fc_use_full_features"):
            setattr(cfgs.pipeline, flag, True)
        cfgs.pipeline.t_embedding_tau = 6
        cfgs.pipeline.activation_function = "relu"
        cfgs.pipeline.mapping_pos_n_freq = 6
        pipes.append(ref.pipelines.load_pipeline(cfgs))
    ours, theirs = pipes
    sd, sd_ref = ours.state_dict(), theirs.state_dict()
    assert list(sd.keys()) == list(sd_ref.keys())
    assert all(tuple(sd[k].shape) == tuple(sd_ref[k].shape) for k in sd)
    assert tuple(sd["model_coarse.semantic_prediction.0.weight"].shape) == (512, 518)
    assert tuple(sd["model_coarse.sky_color.0.weight"].shape) == (512, 3)
    assert tuple(sd["model_coarse.fc_net.0.weight"].shape) == (512, 36) and tuple(sd["model_coarse.fc_net.8.weight"].shape) == (512, 548)
    assert ours.models["coarse"].relu
    assert tuple(ours.models["t"].weight.shape) == tuple(theirs.models["t"].weight.shape) == (cfgs.pipeline.t_embedding_vocab, 6)
    theirs.load_state_dict(sd)
    ours.load_state_dict(sd_ref)


def test_head_variants_build_the_same_state_dict_as_the_reference(ref):
    """every head-input variant the plug-in implements (use_tj_for_s, use_tj_instead_of_beta, use_separate_beta_for_s,
    use_separate_tj_for_semantic) switched on at once: same models (incl. the second embedding "t_s"), same state_dict keys
    and shapes as the reference pipeline built from the same config, checkpoints load both ways"""
    pipes = []
    for dotted in ("semnerf_b200.pipelines.RSSemanticB200Pipeline", "semantic.pipelines.rs_semantic.RSSemanticPipeline"):
        cfgs = _cfgs(ref, "rs_semantic.toml", dotted)
        for flag in ("use_tj_for_s", "use_tj_instead_of_beta", "use_separate_beta_for_s", "use_separate_tj_for_semantic"):
            setattr(cfgs.pipeline, flag, True)
        pipes.append(ref.pipelines.load_pipeline(cfgs))
    ours, theirs = pipes
    assert set(ours.models) == set(theirs.models) == {"coarse", "t", "t_s"}
    sd, sd_ref = ours.state_dict(), theirs.state_dict()
    assert list(sd.keys()) == list(sd_ref.keys())
    assert all(tuple(sd[k].shape) == tuple(sd_ref[k].shape) for k in sd)
    assert tuple(sd["model_coarse.semantic_prediction.0.weight"].shape) == (256, 516)
    assert "model_coarse.semantic_beta_from_xyz.2.weight" in sd and ours.models["coarse"].number_of_outputs == 10 + 6
    ours.load_state_dict(sd_ref)
    theirs.load_state_dict(ours.state_dict())

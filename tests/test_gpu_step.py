"""-m gpu: the training step (fused K3 + loss kernel, direct step, CUDA-graph replay) and the callers either side of the path
(whole-image rendering, point-cloud extraction, the device ray table) against the oracle and the goldens frozen from the
reference (oracle/pin_against_reference.py: SemanticLoss / SemanticCarRegLoss / DepthLoss / SatNerfLoss / SNerfLoss,
batched_inference, get_xyz_from_nerf_prediction, StandardNormalization.denormalize).

Tolerances: bf16 tensor-core MLP against the fp32 oracle - loss terms 1 % relative (they are means of per-ray terms that each
carry ~1e-3 of bf16 noise), gradient cosine >= 0.999 (north_star), rgb / depth 1e-3 at the reference initialisers."""
import os

import numpy as np
import pytest
import torch

from oracle import render_oracle as O
from oracle.step_cases import BATCHED_CASE, BATCHED_CHUNK, CAR, oracle_step_loss
from tests.helpers import GOLDEN
from semnerf_b200.trainer import EMB_PAD
from tests.test_gpu_kernels import DEV, _lib_or_fail, _model

pytestmark = pytest.mark.gpu

TERMS = ("color", "logbeta", "semantic", "car_reg", "sc_term2", "sc_term3", "ds", "semantic_logbeta")


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b) / max(1e-300, float(a.norm() * b.norm()))


def _trainer(kind, C, S, seed, name="", **kw):
    """a Trainer whose parameters are the oracle's deterministic ones (so the oracle can be evaluated on the same weights);
    name: architecture tokens of O.spec_for_case (`full` = fc_use_full_features, `tauN` = t_embedding_tau)"""
    from semnerf_b200.trainer import Trainer, default_cfgs
    spec = O.spec_for_case(name, kind, C)
    cfgs = default_cfgs(kind, n_samples=S, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=2,
                        fc_use_full_features=spec.full_features, t_embedding_tau=spec.tau,
                        activation_function="siren" if spec.siren else "relu", mapping_pos_n_freq=spec.n_freq)
    tr = Trainer(cfgs, kind, C, device=DEV, car_index=CAR, seed=0, **kw)
    params, emb = O.make_params(spec, seed=seed)
    tr.models["coarse"].load_state_dict(params)
    if "t" in tr.models:
        tr.models["t"].weight.data.copy_(emb)
    return tr, spec, params, emb


def _batches(n, nd, C, seed, with_mask):
    """the rgb batch as the reference's semantic dataset delivers it - uint8 (N,1) labels, bool sparsity mask
    (semantic/dataset/semantic_dataset.py:45-87) - and a depth batch (depths (N,1), weights (N,1))"""
    rng = np.random.Generator(np.random.PCG64(seed))
    rays, extras = O.synthetic_rays(n, seed=seed)
    batch = {"rays": rays, "extras": extras, "rgbs": torch.from_numpy(rng.uniform(0, 1, (n, 3))).float()}
    if C:
        lab = torch.from_numpy(rng.integers(0, C, (n, 1))).to(torch.uint8)
        lab[: n // 16] = CAR if CAR < C else 0
        batch["semantic"] = lab
        if with_mask:
            batch["semantic_sparsity_mask"] = torch.from_numpy(rng.uniform(0, 1, n) < 0.6)
    dr, de = O.synthetic_rays(nd, seed=seed + 1)
    depth = {"rays": dr, "extras": de, "depths": torch.from_numpy(rng.uniform(0.1, 0.5, (nd, 1))).float(),
             "weights": torch.from_numpy(rng.uniform(0, 1, (nd, 1))).float()}
    return batch, depth


def _to_dev(b):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()}


@pytest.mark.parametrize("kind,C,epoch,with_depth,with_mask,S,name",
                         [("semantic", 6, 3, True, True, 64, ""), ("semantic", 5, 1, False, False, 64, ""),
                          ("satnerf", 0, 3, True, False, 64, ""), ("snerf", 0, 3, False, False, 64, ""),
                          ("semantic", 6, 3, True, True, 8, ""),
                          # fc_use_full_features / other embedding widths
                          ("semantic", 6, 3, True, True, 64, "full"), ("satnerf", 0, 3, True, False, 64, "full_tau2"),
                          ("semantic", 6, 3, False, True, 64, "tau8"), ("semantic", 6, 3, True, True, 64, "relu"),
                          ("semantic", 6, 3, True, True, 64, "freq4")])
@pytest.mark.parametrize("direct", [True, False])
def test_fused_training_step_matches_the_oracle(kind, C, epoch, with_depth, with_mask, S, name, direct):
    """The fused K3 + loss kernel (through the direct step and through render_loss under autograd) against the ORACLE'S loss
    stack - itself pinned term by term to the reference's loss modules: every term, the total, every parameter gradient, the
    embedding gradient.  The oracle is evaluated on the z_vals the kernel drew (Philox), read back from the step's buffers /
    passed as `u`."""
    _lib_or_fail()
    n, nd = 640, 256
    tr, spec, params, emb = _trainer(kind, C, S, seed=5, name=name, direct=direct)
    batch, depth = _batches(n, nd, C, seed=17, with_mask=with_mask)
    if not direct:   # the autograd path takes the jitter from the caller: share it with the oracle
        g = torch.Generator().manual_seed(3)
        batch["u"], depth["u"] = torch.rand(n, S, generator=g), torch.rand(nd, S, generator=g)
        from semnerf_b200.renderer import B200Renderer
        p = tr.cfgs.pipeline
        sem, car, color = tr._loss_config(epoch)
        tr.gbuf.zero_()
        tr._bind_grads()
        b = _to_dev(batch)
        loss, terms = tr.renderer.render_loss(
            tr.models, b["rays"], b["extras"], b["rgbs"], b.get("semantic") if sem else None, color=color,
            lambda_s=p.lambda_s if sem else 0.0, ignore_index=CAR if sem else -100, lambda_c=p.lambda_c if car else 0.0,
            car_label=CAR, ignore_mask=b.get("semantic_sparsity_mask"), render_options={"u": b["u"]})
        if with_depth:
            d = _to_dev(depth)
            l_d, t_d = tr.renderer.render_loss(tr.models, d["rays"], d["extras"], None, depth=d["depths"][:, 0],
                                               depth_weights=d["weights"].flatten(), lambda_ds=p.ds_lambda,
                                               render_options={"u": d["u"]})
            loss, terms = loss + l_d, terms + t_d
        loss.backward()
    else:
        loss = tr.training_step(_to_dev(batch), epoch=epoch, depth_batch=_to_dev(depth) if with_depth else None)
        terms = tr.last_loss_terms
        batch["z"] = tr._bufs[("rgb", n)].z.cpu()
        if with_depth:
            depth["z"] = tr._bufs[("depth", nd)].z.cpu()
    got = dict(zip(TERMS, terms.cpu().tolist()))
    g_flat = tr.gbuf[EMB_PAD:].cpu()
    g_emb = tr.gbuf[:tr.n_emb].cpu() if tr.n_emb else None

    p2 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    e2 = emb.clone().requires_grad_(True)
    sem = kind == "semantic"
    beta_loss = epoch >= 2 and kind in ("semantic", "satnerf")
    ref, rterms, _ = oracle_step_loss(O, p2, e2, spec, batch, depth, S, ignore_car=True, use_mask=with_mask,
                                      car_reg=sem and epoch >= 2, use_depth=with_depth, beta_loss=beta_loss)
    ref.backward()
    # ---- terms: the oracle reports colour + log-beta + both solar terms as one "color" entry
    color = got["color"] + got["sc_term2"] + got["sc_term3"] + (got["logbeta"] + 1.5 if beta_loss else 0.0)
    assert abs(color - rterms["color"].item()) <= 1e-2 * abs(rterms["color"].item()), (color, rterms["color"].item())
    if sem:
        assert abs(got["semantic"] - rterms["semantic"].item()) <= 1e-2 * rterms["semantic"].item()
        if "car_reg" in rterms:
            assert abs(got["car_reg"] - rterms["car_reg"].item()) <= 2e-2 * rterms["car_reg"].item() + 1e-6
        else:
            assert got["car_reg"] == 0.0
    if with_depth:
        assert abs(got["ds"] - rterms["ds"].item()) <= 1e-2 * rterms["ds"].item()
    assert abs(loss.item() - ref.item()) <= 1e-2 * abs(ref.item()), (loss.item(), ref.item())
    # ---- gradients
    g_ref = torch.cat([p2[k].grad.flatten() if p2[k].grad is not None else torch.zeros(p2[k].numel()) for k in p2])
    assert _cos(g_flat, g_ref) >= 0.999, _cos(g_flat, g_ref)
    rel = (g_flat.norm() / g_ref.norm()).item()
    assert abs(rel - 1.0) <= 2e-2, rel
    off = 0
    for k in p2:   # every tensor with a non-trivial gradient on its own
        m = p2[k].numel()
        if p2[k].grad is not None and p2[k].grad.norm() > 1e-6 * g_ref.norm():
            # (a ReLU net is noisier per tensor in bf16 than the SIREN models: the [h > 0] masks of near-zero pre-activations
            # flip under rounding; the whole-gradient bar above is the same - see test_nerf_model_and_render_gradients_match_oracle)
            assert _cos(g_flat[off:off + m], p2[k].grad) >= (0.995 if spec.siren else 0.98), k
        off += m
    if g_emb is not None:
        if beta_loss:
            assert _cos(g_emb, e2.grad) >= 0.995
        else:   # the embedding only feeds the uncertainty head, which the beta-free colour loss does not touch
            assert e2.grad.abs().max() == 0 and g_emb.abs().max() == 0


@pytest.mark.parametrize("mode,detach,bs,tj", [("direct", False, False, False), ("direct", True, False, False),
                                               ("autograd", False, False, False), ("module", False, False, False),
                                               ("direct", False, True, False), ("direct", True, True, False),
                                               ("autograd", False, True, False), ("module", False, True, False),
                                               ("direct", False, True, True)])
def test_uncertainty_weighted_semantic_loss_matches_the_oracle(mode, detach, bs, tj):
    """`use_beta_for_s` (SemanticUncertaintyLoss, semantic/components/loss.py:6-32,68-114): lambda_s * CE_mean * mean_r 1/(2 beta_r^2),
    a product of two batch means - the fused kernel takes them from a statistics pre-pass.  All three trainer paths against the
    oracle's loss (pinned to the reference's module by oracle/pin_against_reference.py), with and without detach_beta_for_s.
    bs: `use_separate_beta_for_s` - the semantic head's own uncertainty head (output column 9, rs_semantic.py:228-237) takes
    the place of beta in this loss and adds its log term.  tj: all head variants at once (`use_tj_for_s` on top: the embedding
    then feeds the uncertainty, the semantic-uncertainty and the semantic head)."""
    from semnerf_b200.trainer import Trainer, default_cfgs
    _lib_or_fail()
    C, S, n = 6, 64, 640
    cfgs = default_cfgs("semantic", n_samples=S, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=2, use_beta_for_s=True,
                        detach_beta_for_s=detach, lambda_s=0.4,   # a larger weight: the term must matter in the gradient
                        use_separate_beta_for_s=bs, use_tj_for_s=tj)
    kw = {"direct": dict(direct=True), "autograd": dict(direct=False), "module": dict(fused_loss=False)}[mode]
    tr = Trainer(cfgs, "semantic", C, device=DEV, car_index=CAR, seed=0, **kw)
    spec = O.ModelSpec(kind="semantic", n_classes=C, separate_beta_s=bs, tj_for_s=tj)
    params, emb = O.make_params(spec, seed=5)
    tr.models["coarse"].load_state_dict(params)
    tr.models["t"].weight.data.copy_(emb)
    assert tr.models["coarse"].number_of_outputs == 9 + C + (1 if bs else 0)
    batch, depth = _batches(n, 64, C, seed=23, with_mask=True)
    assert tr._sem_unc(3) == (2 if detach else 1) and tr._sem_unc(1) == 0      # plain cross-entropy before first_beta_epoch
    loss = tr.training_step(_to_dev(batch), epoch=3)
    # the jitter every path drew: Philox keyed on (step 1, ray) - recover z from a render with the same key
    with torch.no_grad():
        z = tr.renderer.render_rays(tr.models, batch["rays"].to(DEV), batch["extras"].to(DEV),
                                    render_options={"seed": 1, "ray_offset": 0, "solar_pass": False})["_z_vals_coarse"].cpu()
    if mode == "direct":
        assert torch.equal(z, tr._bufs[("rgb", n)].z.cpu())
    batch["z"] = z
    p2 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    e2 = emb.clone().requires_grad_(True)
    ref, rterms, _ = oracle_step_loss(O, p2, e2, spec, batch, depth, S, ignore_car=True, use_mask=True, car_reg=True,
                                      use_depth=False, beta_loss=True, lambda_s=0.4, sem_unc=2 if detach else 1)
    ref.backward()
    if mode != "module":
        got = dict(zip(TERMS, tr.last_loss_terms.cpu().tolist()))
        mine = got["semantic"] + (got["semantic_logbeta"] + 1.5 * 0.4 if bs else 0.0)     # the oracle reports both as one term
        assert abs(mine - rterms["semantic"].item()) <= 1e-2 * rterms["semantic"].item(), (mine, rterms["semantic"].item())
    else:
        mine = float(tr.last_loss_dict["coarse_semantic"].detach()) + \
            (float(tr.last_loss_dict["coarse_semantic_logbeta"].detach()) if bs else 0.0)
        assert abs(mine - rterms["semantic"].item()) <= 1e-2 * rterms["semantic"].item()
    assert abs(loss.item() - ref.item()) <= 1e-2 * abs(ref.item())
    # NB: the parameters have moved (Adam ran), the gradient buffer still holds this step's gradient
    g_ref = torch.cat([p2[k].grad.flatten() if p2[k].grad is not None else torch.zeros(p2[k].numel()) for k in p2])
    assert _cos(tr.gbuf[EMB_PAD:].cpu(), g_ref) >= 0.999
    assert _cos(tr.gbuf[:tr.n_emb].cpu(), e2.grad) >= 0.995
    # the uncertainty heads' own gradients are where detach / no detach differ most
    for name in ["beta_from_xyz.2.weight"] + (["semantic_beta_from_xyz.2.weight", "semantic_beta_from_xyz.0.weight"] if bs else []):
        off, m = tr.models["coarse"].offset_of(name), p2[name].numel()
        got_g, ref_g = tr.gbuf[EMB_PAD + off: EMB_PAD + off + m].cpu(), p2[name].grad
        if ref_g is None or ref_g.abs().max() == 0:     # detached: the head receives nothing from this loss
            assert got_g.abs().max() <= 1e-7 * g_ref.abs().max()
        else:
            assert _cos(got_g, ref_g) >= 0.995, name
    if bs and mode != "module":
        got = dict(zip(TERMS, tr.last_loss_terms.cpu().tolist()))
        assert got["semantic_logbeta"] != 0.0


def test_label_dtypes_and_counts_are_device_side():
    """uint8 (N,1) labels (what the reference's dataset yields) and int64 (N,) labels give the same step; the masked-mean
    denominators come from snb_label_counts with the loss kernel's own predicates; out-of-range labels are reported."""
    from semnerf_b200.autograd import as_labels, as_ray_mask, label_counts
    from semnerf_b200.raytable import DeviceRayTable
    _lib_or_fail()
    n, C = 1000, 6
    g = torch.Generator().manual_seed(0)
    lab = torch.randint(0, C, (n, 1), generator=g).to(torch.uint8)
    mask = torch.rand(n, generator=g) < 0.5
    c = label_counts(as_labels(lab.to(DEV)), as_ray_mask(mask.to(DEV)), C, CAR, CAR).cpu()
    flat = lab.flatten().long()
    assert c[0].item() == int((mask & (flat != CAR)).sum()) and c[1].item() == int((mask & (flat == CAR)).sum())
    assert c[2].item() == 0
    bad = flat.clone()
    bad[:7] = 9
    assert label_counts(bad.to(DEV), None, C, CAR, CAR)[2].item() == 7
    with pytest.raises(ValueError):
        DeviceRayTable({"rays": torch.zeros(n, 8), "semantic": bad}, device=DEV).validate_labels(C)
    DeviceRayTable({"rays": torch.zeros(n, 8), "semantic": lab}, device=DEV).validate_labels(C)
    losses = []
    for labels in (lab, flat):
        tr, *_ = _trainer("semantic", C, 8, seed=2)
        batch, _ = _batches(n, 8, C, seed=4, with_mask=False)
        batch["semantic"] = labels
        losses.append(tr.training_step(_to_dev(batch), epoch=3).item())
    assert abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[0])    # same kernels, same data (atomic-add order only)


def test_cuda_graph_replay_equals_the_eager_direct_step():
    """graph=True captures the whole step (K1 .. Adam + re-pack) once per configuration and replays it; the Philox key and the
    Adam step counter are read from device memory, so every replay is a fresh step: same losses and parameters as the eager
    direct step (up to the order of the fp32 split-K reduce-adds)."""
    _lib_or_fail()
    n, nd, C, S = 1024, 512, 6, 64
    batches = [_batches(n, nd, C, seed=30 + i, with_mask=True) for i in range(3)]
    out = {}
    for graph in (False, True):
        tr, *_ = _trainer("semantic", C, S, seed=9, graph=graph)
        losses = []
        for i in range(6):
            b, d = batches[i % 3]
            losses.append(tr.training_step(_to_dev(b), epoch=3, depth_batch=_to_dev(d) if i < 4 else None).item())
        if graph:
            assert sum(1 for e in tr._graphs.values() if e[0] == "graph") == 2    # with and without the depth batch
        out[graph] = (losses, tr.pbuf.clone())
    le, pe = out[False]
    lg, pg = out[True]
    assert np.allclose(le, lg, rtol=2e-4), (le, lg)
    # Adam normalises every gradient entry by its own magnitude, so the few entries whose gradient is fp32 reduce-order noise
    # may move differently; everything else is the same to ~1e-6
    assert (pe - pg).abs().float().quantile(0.999).item() <= 1e-4 and _cos(pe, pg) >= 0.999999
    assert le[-1] < le[0]


# ---------------------------------------------------------------------------------------------------------------------
# a11: whole-image rendering in chunks (batched_inference / BaseRayPipeline.forward)
# ---------------------------------------------------------------------------------------------------------------------
def test_render_image_equals_the_reference_batched_inference_golden():
    """Trainer.render_image over 2.5 chunks against what the reference's own batched_inference returned for the same rays,
    weights and jitter (tests/golden/batched_sem_c6_s8.npz), and against one un-chunked render_rays call."""
    _lib_or_fail()
    name, kind, C, feat, n, s, sc, seed = BATCHED_CASE
    gold = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    tr, spec, params, emb = _trainer(kind, C, s, seed=seed)
    rays, extras = O.synthetic_rays(n, seed=seed)
    u = torch.from_numpy(gold["u"]).to(DEV)
    keys = ("rgb_coarse", "depth_coarse", "semantic_label_coarse", "weights_coarse", "sun_sc_coarse")
    img = tr.render_image(rays.to(DEV), extras.to(DEV), chunk=BATCHED_CHUNK, keys=keys, u=u)
    with torch.no_grad():
        one = tr.renderer.render_rays(tr.models, rays.to(DEV), extras.to(DEV), render_options={"u": u})
    for k in keys:
        assert img[k].shape == gold[k].shape, k
        assert torch.equal(img[k], one[k]), k                      # chunking changes nothing, bit for bit
    assert np.abs(img["rgb_coarse"].cpu().numpy() - gold["rgb_coarse"]).max() <= 1e-3
    assert np.abs(img["depth_coarse"].cpu().numpy() - gold["depth_coarse"]).max() <= 1e-3
    assert np.abs(img["weights_coarse"].cpu().numpy() - gold["weights_coarse"]).max() <= 2e-3
    assert np.abs(img["sun_sc_coarse"].cpu().numpy() - gold["sun_sc_coarse"]).max() <= 2e-3


def test_render_image_is_invariant_to_chunking_and_sharding_under_philox():
    """in-kernel Philox jitter keyed on (seed, global ray index): any chunk size and any split of the image across ranks
    (contiguous ray ranges with their ray_offset) gives the same image, bit for bit; and it agrees with the oracle evaluated
    on the same z_vals."""
    _lib_or_fail()
    C, S, n = 6, 64, 2500
    tr, spec, params, emb = _trainer("semantic", C, S, seed=11)
    rays, extras = O.synthetic_rays(n, seed=3)
    r, e = rays.to(DEV), extras.to(DEV)
    keys = ("rgb_coarse", "depth_coarse", "semantic_label_coarse", "semantic_logits_coarse", "_z_vals_coarse")
    a = tr.render_image(r, e, chunk=1000, keys=keys, seed=7)          # 2.5 chunks
    b = tr.render_image(r, e, chunk=4096, keys=keys, seed=7)          # one chunk
    halves = [tr.render_image(r[lo:hi], e[lo:hi], chunk=700, keys=keys, seed=7, ray_offset=lo) for lo, hi in ((0, 1250), (1250, n))]
    for k in keys:
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(torch.cat([h[k] for h in halves]), a[k]), k
    c = tr.render_image(r, e, chunk=1000, keys=keys, seed=8)
    assert not torch.equal(a["_z_vals_coarse"], c["_z_vals_coarse"])   # another seed, another jitter
    with torch.no_grad():
        ref = O.render_rays(params, emb, spec, rays, extras, S, z=a["_z_vals_coarse"].cpu(), sc_lambda=0.0)
    assert (a["rgb_coarse"].cpu() - ref["rgb_coarse"]).abs().max() <= 1e-3
    assert (a["depth_coarse"].cpu() - ref["depth_coarse"]).abs().max() <= 1e-3
    margin = ref["semantic_logits_coarse"].topk(2, -1)[0]
    clear = (margin[:, 0] - margin[:, 1]) > 5e-3
    assert (a["semantic_label_coarse"].cpu() == ref["semantic_label_coarse"])[clear].float().mean().item() >= 0.999


# ---------------------------------------------------------------------------------------------------------------------
# 8f rank 3: streaming depth / point-cloud extraction
# ---------------------------------------------------------------------------------------------------------------------
def test_extract_pointcloud_against_the_reference_chain_golden():
    """extract_pointcloud (depth-only / depth + rgb head masks, chunked, device float64 xyz + de-normalisation) against the
    golden produced by the reference's batched_inference -> get_xyz_from_nerf_prediction -> StandardNormalization.denormalize."""
    from semnerf_b200.pointcloud import extract_pointcloud
    _lib_or_fail()
    name, kind, C, feat, n, s, sc, seed = BATCHED_CASE
    gold = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    tr, spec, params, emb = _trainer(kind, C, s, seed=seed)
    rays, extras = O.synthetic_rays(n, seed=seed)
    u = torch.from_numpy(gold["u"]).to(DEV)
    scale = float(gold["scale"])
    for want_rgb in (True, False):
        pc = extract_pointcloud(tr.renderer, tr.models, rays.to(DEV), extras.to(DEV), center=gold["center"].tolist(),
                                scale=scale, chunk=BATCHED_CHUNK, want_rgb=want_rgb, u=u)
        assert pc["xyz"].dtype == torch.float64 and pc["xyz"].shape == (n, 3)
        assert np.abs(pc["depth"].cpu().numpy() - gold["depth_coarse"]).max() <= 1e-3
        # the xyz chain itself is exact: applied to OUR depth it reproduces the oracle's formulas bit for bit ...
        mine = O.denormalize(O.xyz_from_depth(rays, pc["depth"].cpu()), gold["center"].tolist(), scale)
        assert torch.equal(pc["xyz"].cpu(), mine) and torch.equal(pc["xyz_n"].cpu(), O.xyz_from_depth(rays, pc["depth"].cpu()))
        # ... and lands within the depth tolerance (scaled to scene units) of the reference's points
        assert np.abs(pc["xyz_n"].cpu().numpy() - gold["xyz_n"]).max() <= 1e-3
        assert np.abs(pc["xyz"].cpu().numpy() - gold["xyz"]).max() <= 1e-3 * scale
        if want_rgb:
            assert np.abs(pc["rgb"].cpu().numpy() - gold["rgb_coarse"]).max() <= 1e-3
    # the depth-only head mask evaluates the same trunk + sigma as the full pass
    full = extract_pointcloud(tr.renderer, tr.models, rays.to(DEV), extras.to(DEV), chunk=7, want_rgb=True, u=u)
    assert (full["depth"] - pc["depth"]).abs().max().item() <= 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# 8f rank 2: the device-resident ray table drives the trainer
# ---------------------------------------------------------------------------------------------------------------------
def test_device_ray_table_epoch_drives_the_trainer_on_cuda():
    from semnerf_b200 import synth
    from semnerf_b200.raytable import DeviceRayTable, ZippedTables
    _lib_or_fail()
    C, S, n, nd = 6, 16, 5000, 1200
    rays, extras = synth.make_rays(n, seed=0)
    rgbs, labels, _ = synth.make_targets(rays, C, seed=0)
    dr, de = synth.make_rays(nd, seed=1)
    _, _, depths = synth.make_targets(dr, C, seed=1)
    rgb_t = DeviceRayTable({"rays": rays, "extras": extras, "rgbs": rgbs, "semantic": labels.to(torch.uint8).view(-1, 1),
                            "semantic_sparsity_mask": torch.ones(n, dtype=torch.bool), "ids": torch.arange(n)}, device=DEV)
    dep_t = DeviceRayTable({"rays": dr, "extras": de, "depths": depths.view(-1, 1), "weights": torch.ones(nd, 1)}, device=DEV)
    rgb_t.validate_labels(C)
    assert all(v.is_cuda for v in rgb_t.tensors.values())
    tr, *_ = _trainer("semantic", C, S, seed=3)
    zipped = ZippedTables({"rgb": rgb_t, "depth": dep_t}, {"rgb": 1024, "depth": 512}, seed=5)
    seen, losses = [], []
    for step, b in enumerate(zipped.epoch(0)):
        assert b["rgb"]["rays"].is_cuda and b["rgb"]["_global_rays"] == b["rgb"]["rays"].shape[0]
        assert torch.equal(b["rgb"]["rays"], rgb_t.tensors["rays"][b["rgb"]["ids"]])      # keys stay aligned on the device
        seen.append(b["rgb"]["ids"])
        losses.append(tr.training_step(b["rgb"], epoch=3, depth_batch=b["depth"] if step < 2 else None))
    assert len(seen) == 5 and [len(s) for s in seen] == [1024] * 4 + [904]               # drop_last=False: short last batch
    assert torch.equal(torch.sort(torch.cat(seen)).values.cpu(), torch.arange(n))       # one pass over every ray
    losses = [l.item() for l in losses]
    assert all(np.isfinite(losses))
    # a second epoch (a fresh order) keeps learning
    for step, b in enumerate(zipped.epoch(1)):
        last = tr.training_step(b["rgb"], epoch=3).item()
    assert np.isfinite(last) and last <= losses[2] * 1.05

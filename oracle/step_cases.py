"""Inputs and the oracle's loss stack of one whole training step (RSSemanticTrainingStep.training_step,
semantic/components/training_step.py:12-99).  TEST INFRASTRUCTURE - see ``oracle/__init__.py``: shared by
``oracle/pin_against_reference.py`` (which pins it against the reference's loss modules and freezes
``tests/golden/step_*.npz``) and by the tests."""
from __future__ import annotations

import numpy as np
import torch

from . import render_oracle as O

# name, C, n_rays, n_samples, seed, ignore_car_index, with sparsity mask, car regularisation, depth batch, beta loss
STEP_CASES = [
    ("step_sem_c6_full", 6, 48, 16, 21, True, True, True, True, True),
    ("step_sem_c5_early", 5, 32, 8, 22, False, False, False, True, False),
]
CAR = 4
# a11: the reference's batched_inference over 25 rays in chunks of 10 (2.5 chunks) - (name, kind, C, feat, n, S, sc_lambda, seed)
BATCHED_CASE = ("batched_sem_c6_s8", "semantic", 6, 512, 25, 8, 0.05, 31)
BATCHED_CHUNK = 10


def step_inputs(name, C, n, s, seed, *_):
    """inputs of one training step: the rgb batch (rays, extras, rgbs, uint8 labels (N,1), sparsity mask) and the depth
    batch (rays, extras, depths, weights) - dtypes and shapes as the reference's datasets deliver them
    (semantic/dataset/semantic_dataset.py:45-87, baseline/dataset/satnerf_dataset.py:95-133)."""
    spec = O.ModelSpec(kind="semantic", n_classes=C)
    params, emb = O.make_params(spec, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 500))
    rays, extras = O.synthetic_rays(n, seed=seed)
    d_rays, d_extras = O.synthetic_rays(n // 2, seed=seed + 1)
    batch = {
        "rays": rays, "extras": extras,
        "rgbs": torch.from_numpy(rng.uniform(0, 1, (n, 3))).float(),
        "semantic": torch.from_numpy(rng.integers(0, C, (n, 1))).to(torch.uint8),
        "semantic_sparsity_mask": torch.from_numpy(rng.uniform(0, 1, n) < 0.7),
        "u": torch.from_numpy(rng.uniform(0, 1, (n, s))).float(),
    }
    batch["semantic"][:3] = CAR if CAR < C else 0            # a few car rays, some of them masked out below
    batch["semantic_sparsity_mask"][1] = False
    batch["semantic_sparsity_mask"][0] = True
    depth = {
        "rays": d_rays, "extras": d_extras,
        "depths": torch.from_numpy(rng.uniform(0.1, 0.5, (n // 2, 1))).float(),
        "weights": torch.from_numpy(rng.uniform(0, 1, (n // 2, 1))).float(),
        "u": torch.from_numpy(rng.uniform(0, 1, (n // 2, s))).float(),
    }
    return spec, params, emb, batch, depth


def oracle_step_loss(O_, params, emb, spec, batch, depth, s, ignore_car, use_mask, car_reg, use_depth, beta_loss,
                     lambda_s=0.04, lambda_c=0.1, ds_lambda=1000.0, sc_lambda=0.05, sem_unc=0):
    """the oracle's restatement of RSSemanticTrainingStep.training_step (semantic/components/training_step.py:12-99)"""
    res = O_.render_rays(params, emb, spec, batch["rays"], batch["extras"], s, u=batch.get("u"), z=batch.get("z"),
                         sc_lambda=sc_lambda)
    terms = {}
    terms["color"] = (O_.satnerf_loss if beta_loss else O_.snerf_loss)(res, batch["rgbs"], lambda_sc=sc_lambda)
    if use_depth:
        tmp = O_.render_rays(params, emb, spec, depth["rays"], depth["extras"], s, u=depth.get("u"), z=depth.get("z"),
                             sc_lambda=sc_lambda)
        terms["ds"] = O_.depth_loss(tmp, torch.flatten(depth["depths"][:, 0]), torch.flatten(depth["weights"]), ds_lambda)
    mask = batch["semantic_sparsity_mask"] if use_mask else None
    if spec.kind != "semantic":     # the baseline pipelines' step: colour (+ depth) only (baseline/components/training_step.py)
        return sum(terms.values()), terms, res
    if sem_unc:   # use_beta_for_s (training_step.py:66-75): 1 = SemanticUncertaintyLoss, 2 = with detach_beta_for_s
        terms["semantic"] = O_.semantic_uncertainty_loss(res, batch["semantic"], lambda_s, CAR if ignore_car else -100, mask,
                                                         detach_beta=(sem_unc == 2))
    else:
        terms["semantic"] = O_.semantic_loss(res, batch["semantic"], lambda_s, CAR if ignore_car else -100, mask)
    if car_reg:
        terms["car_reg"] = O_.car_reg_loss(res, batch["semantic"], CAR, lambda_c, mask)
    return sum(terms.values()), terms, res

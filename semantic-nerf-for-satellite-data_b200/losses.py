"""Loss modules that sit directly on the render path's outputs - same class names, constructor
arguments, input keys and returned (loss, loss_dict) as the reference's
baseline/components/loss.py and semantic/components/loss.py, so a training step written against the
reference runs unchanged.  Plain PyTorch on the device (adjacent to the hot
path, not part of it)."""
from __future__ import annotations

import torch


def solar_correction(loss_dict, inputs, typ, lambda_sc=0.05):
    """baseline/components/loss.py:4-13"""
    sun_sc = inputs[f"sun_sc_{typ}"].squeeze(-1)
    term2 = torch.sum(torch.square(inputs[f"transparency_sc_{typ}"].detach() - sun_sc), -1)
    term3 = 1 - torch.sum(inputs[f"weights_sc_{typ}"].detach() * sun_sc, -1)
    loss_dict[f"{typ}_sc_term2"] = lambda_sc / 3.0 * torch.mean(term2)
    loss_dict[f"{typ}_sc_term3"] = lambda_sc / 3.0 * torch.mean(term3)
    return loss_dict


def uncertainty_aware_loss(loss_dict, inputs, gt_rgb, typ, beta_min=0.05):
    """baseline/components/loss.py:16-27"""
    beta = torch.sum(inputs[f"weights_{typ}"].unsqueeze(-1) * inputs["beta_coarse"], -2) + beta_min
    loss_dict[f"{typ}_color"] = ((inputs[f"rgb_{typ}"] - gt_rgb) ** 2 / (2 * beta ** 2)).mean()
    loss_dict[f"{typ}_logbeta"] = (3 + torch.log(beta).mean()) / 2
    return loss_dict


class DepthLoss(torch.nn.Module):
    """baseline/components/loss.py:30-47"""

    def __init__(self, lambda_ds=1.0):
        super().__init__()
        self.lambda_ds = lambda_ds / 3.0

    def forward(self, inputs, targets, weights=1.0):
        d = {"coarse_ds": self.lambda_ds * torch.mean(weights * (inputs["depth_coarse"] - targets) ** 2)}
        return sum(d.values()), d


class SatNerfLoss(torch.nn.Module):
    """baseline/components/loss.py:50-68"""

    def __init__(self, lambda_sc=0.0, solar_correction_enabled=True):
        super().__init__()
        self.lambda_sc, self.solar_correction_enabled = lambda_sc, solar_correction_enabled

    def forward(self, inputs, targets):
        d = uncertainty_aware_loss({}, inputs, targets, "coarse")
        if self.lambda_sc > 0 and self.solar_correction_enabled:
            d = solar_correction(d, inputs, "coarse", self.lambda_sc)
        return sum(d.values()), d


class SNerfLoss(torch.nn.Module):
    """baseline/components/loss.py:71-94"""

    def __init__(self, lambda_sc=0.05, solar_correction_enabled=True):
        super().__init__()
        self.lambda_sc, self.solar_correction_enabled = lambda_sc, solar_correction_enabled

    def forward(self, inputs, targets):
        d = {"coarse_color": torch.nn.functional.mse_loss(inputs["rgb_coarse"], targets)}
        if self.lambda_sc > 0 and self.solar_correction_enabled:
            d = solar_correction(d, inputs, "coarse", self.lambda_sc)
        return sum(d.values()), d


class NerfLoss(torch.nn.Module):
    """baseline/components/loss.py:97-110 (coarse network only: n_importance = 0)"""

    def forward(self, inputs, targets):
        d = {"coarse_color": torch.nn.functional.mse_loss(inputs["rgb_coarse"], targets)}
        return sum(d.values()), d


def _masked(logits, targets, ignore_mask):
    """`inputs[...][ignore_mask], targets[ignore_mask].squeeze()` (semantic/components/loss.py:52-55); labels arrive as the
    dataset delivers them - uint8, (N,1) (framework/util/img_utils.py::load_tensor_from_cls_geotiff)"""
    tgt = targets.reshape(-1).long()
    if ignore_mask is not None:
        m = ignore_mask.reshape(-1).bool()
        logits, tgt = logits[m], tgt[m]
    return logits, tgt


class SemanticLoss(torch.nn.Module):
    """semantic/components/loss.py:35-65"""

    def __init__(self, lambda_s, car_index, ignore_car_index=False):
        super().__init__()
        self.lambda_s = lambda_s
        self.loss = torch.nn.CrossEntropyLoss(ignore_index=car_index if ignore_car_index else -100)

    def forward(self, inputs, targets, ignore_mask=None):
        logits, tgt = _masked(inputs["semantic_logits_coarse"], targets, ignore_mask)
        d = {"coarse_semantic": self.lambda_s * self.loss(logits, tgt)}
        return sum(d.values()), d


class SemanticUncertaintyLoss(torch.nn.Module):
    """semantic/components/loss.py:6-32,68-114 (`use_beta_for_s`): the cross-entropy mean weighted per ray by 1 / (2 beta^2)
    of the composited uncertainty - of the separate semantic uncertainty head when the render returned
    `beta_semantic_coarse`, which also adds its own log-beta term."""

    def __init__(self, lambda_s, car_index, detach_beta_for_s=False, ignore_car_index=False, beta_min=0.05):
        super().__init__()
        self.lambda_s, self.detach_beta_for_s, self.beta_min = lambda_s, detach_beta_for_s, beta_min
        self.cross_entropy = torch.nn.CrossEntropyLoss(ignore_index=car_index if ignore_car_index else -100)

    def forward(self, inputs, targets, ignore_mask=None):
        beta_in = inputs.get("beta_semantic_coarse", inputs["beta_coarse"])
        if self.detach_beta_for_s:
            beta_in = beta_in.detach().clone()
        beta = torch.sum(inputs["weights_coarse"].unsqueeze(-1) * beta_in, -2) + self.beta_min
        logits, tgt = _masked(inputs["semantic_logits_coarse"], targets, ignore_mask)
        d = {"coarse_semantic": self.lambda_s * (self.cross_entropy(logits, tgt) / (2 * beta ** 2)).mean()}
        if "beta_semantic_coarse" in inputs:
            d["coarse_semantic_logbeta"] = self.lambda_s * (3 + torch.log(beta).mean()) / 2
        return sum(d.values()), d


class SemanticCarRegLoss(torch.nn.Module):
    """semantic/components/loss.py:117-157: push the composited uncertainty of car-labelled rays to 1
    (the 'transient regularisation')."""

    def __init__(self, lambda_c, car_label):
        super().__init__()
        self.lambda_c, self.car_label = lambda_c, car_label

    def forward(self, inputs, targets, ignore_mask=None):
        unc = torch.sum(inputs["weights_coarse"].unsqueeze(-1) * inputs["beta_coarse"], -2)
        mask = targets.reshape(-1) == self.car_label
        if ignore_mask is not None:
            mask = mask & ignore_mask.reshape(-1).bool()
        # masked mean without a data-dependent shape (no host sync): mean over selected rays.  Deviation, on purpose: with NO
        # selected ray the reference's MSELoss of an empty tensor is NaN (and poisons the step); this term is then 0.
        sq = (1.0 - unc.reshape(-1)) ** 2 * mask
        d = {"coarse_car_reg_loss": self.lambda_c * sq.sum() / mask.sum().clamp_min(1)}
        return sum(d.values()), d

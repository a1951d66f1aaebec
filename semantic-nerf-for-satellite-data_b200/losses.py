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


class SemanticLoss(torch.nn.Module):
    """semantic/components/loss.py:35-65"""

    def __init__(self, lambda_s, car_index, ignore_car_index=False):
        super().__init__()
        self.lambda_s = lambda_s
        self.loss = torch.nn.CrossEntropyLoss(ignore_index=car_index if ignore_car_index else -100)

    def forward(self, inputs, targets, ignore_mask=None):
        logits, tgt = inputs["semantic_logits_coarse"], targets.reshape(-1)
        if ignore_mask is not None:
            logits, tgt = logits[ignore_mask], tgt[ignore_mask]
        d = {"coarse_semantic": self.lambda_s * self.loss(logits, tgt)}
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
            mask = mask & ignore_mask.reshape(-1)
        # masked mean without a data-dependent shape (no host sync): mean over selected rays
        sq = (1.0 - unc.reshape(-1)) ** 2 * mask
        d = {"coarse_car_reg_loss": self.lambda_c * sq.sum() / mask.sum().clamp_min(1)}
        return sum(d.values()), d

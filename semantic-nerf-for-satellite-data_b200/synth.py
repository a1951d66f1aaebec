"""Synthetic satellite-like rays for benchmarks and smoke tests (DFC2019 is not available offline).
Value ranges follow the reference's data layout (framework/components/rays.py:7-64) and scene
normalisation (baseline/components/normalization.py:60-79); see SURVEY section 8d."""
from __future__ import annotations

import numpy as np
import torch


def make_rays(n_rays: int, seed: int = 0, n_images: int = 17):
    """rays (N,8) [o d near far], extras (N,4) [sun_d ts] as float32 CPU tensors."""
    rng = np.random.Generator(np.random.PCG64(seed))
    o = np.concatenate([rng.uniform(-1, 1, (n_rays, 2)), rng.uniform(0.15, 0.35, (n_rays, 1))], 1)
    d = np.array([0.10, 0.05, -1.0]) + 0.02 * rng.standard_normal((n_rays, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    far = rng.uniform(0.45, 0.65, (n_rays, 1))
    el, az = np.deg2rad(rng.uniform(30, 70, n_images)), np.deg2rad(rng.uniform(100, 200, n_images))
    sun = np.stack([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)], 1)
    ts = rng.integers(0, n_images, n_rays)
    rays = np.concatenate([o, d, np.zeros((n_rays, 1)), far], 1).astype(np.float32)
    extras = np.concatenate([sun[ts], ts[:, None]], 1).astype(np.float32)
    return torch.from_numpy(rays), torch.from_numpy(extras)


def make_targets(rays: torch.Tensor, n_classes: int, seed: int = 0):
    """procedural scene: colour and class are smooth functions of the ray origin, so a few hundred
    optimisation steps give a model with real class margins."""
    x, y = rays[:, 0], rays[:, 1]
    rgb = torch.stack([0.5 + 0.4 * torch.sin(3 * x), 0.5 + 0.4 * torch.cos(2 * y), 0.5 + 0.3 * torch.sin(2 * x + y)], 1)
    if n_classes > 0:
        label = ((x + 1) * 0.5 * n_classes).long().clamp(0, n_classes - 1)
    else:
        label = torch.zeros(rays.shape[0], dtype=torch.long)
    g = torch.Generator().manual_seed(seed)
    depth = 0.2 + 0.2 * torch.rand(rays.shape[0], generator=g)
    return rgb.float().clamp(0, 1), label, depth

"""Streaming depth / point-cloud extraction: the caller side of the dense eval sweep.

Replaces `eval/extract_pointcloud.py:66-94` + `batched_inference` (eval/utils/util.py:13-42) +
`get_xyz_from_nerf_prediction` (baseline/dataset/satnerf_dataset.py:156-171) + `StandardNormalization.denormalize`
(baseline/components/normalization.py:44-58): rays are rendered chunk by chunk with the depth-only head mask
(trunk + sigma: everything the sweep consumes besides colour), per-chunk results land in preallocated `(H*W, .)`
buffers (the reference `torch.cat`s a growing dict per chunk and materialises ~2.3 GB of per-sample tensors per
image), and `xyz = o + d * depth` plus the de-normalisation `xyz * range + center` run on the device in float64 as
the reference does on the host.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch


@torch.no_grad()
def render_depth_rgb(renderer, models, rays: torch.Tensor, extras: torch.Tensor, chunk: int = 40960,
                     want_rgb: bool = True, seed: int = 1, ray_offset: int = 0, u=None) -> Dict[str, torch.Tensor]:
    """depth (N,) [and rgb (N,3)] of a whole view, chunked; `want_rgb=False` evaluates trunk + sigma only.
    The jitter is keyed on (seed, global ray index) - independent of the chunking; `u` (N,S) replaces it (parity tests)."""
    n = rays.shape[0]
    depth = torch.empty(n, dtype=torch.float32, device=rays.device)
    rgb = torch.empty(n, 3, dtype=torch.float32, device=rays.device) if want_rgb else None
    for i in range(0, n, chunk):
        opts = {"seed": seed, "ray_offset": ray_offset + i, "heads": "all" if want_rgb else "depth", "solar_pass": False}
        if u is not None:
            opts["u"] = u[i:i + chunk]
        res = renderer.render_rays(models, rays[i:i + chunk], extras[i:i + chunk] if extras is not None else None,
                                   render_options=opts)
        depth[i:i + chunk] = res["depth_coarse"]
        if want_rgb:
            rgb[i:i + chunk] = res["rgb_coarse"]
    out = {"depth": depth}
    if want_rgb:
        out["rgb"] = rgb
    return out


def xyz_from_depth(rays: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """normalised end points of the rays at the predicted depth, float64 (satnerf_dataset.py:156-171)."""
    r = rays.double()
    return r[:, 0:3] + r[:, 3:6] * depth.double().view(-1, 1)


def denormalize(xyz_n: torch.Tensor, center: Sequence[float], scale: float) -> torch.Tensor:
    """StandardNormalization.denormalize (normalization.py:44-58): xyz * range + center."""
    c = torch.as_tensor(center, dtype=xyz_n.dtype, device=xyz_n.device)
    return xyz_n * float(scale) + c


@torch.no_grad()
def extract_pointcloud(renderer, models, rays, extras, center: Optional[Sequence[float]] = None, scale: float = 1.0,
                       chunk: int = 40960, want_rgb: bool = True, u=None) -> Dict[str, torch.Tensor]:
    """One view -> {"xyz_n" (N,3) f64 normalised, "xyz" (N,3) f64 scene coordinates, "depth", ["rgb"]}."""
    res = render_depth_rgb(renderer, models, rays, extras, chunk, want_rgb, u=u)
    xyz_n = xyz_from_depth(rays, res["depth"])
    res["xyz_n"] = xyz_n
    res["xyz"] = denormalize(xyz_n, center, scale) if center is not None else xyz_n
    return res

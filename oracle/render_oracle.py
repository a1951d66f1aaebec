"""Plain-PyTorch (CPU, fp32/fp64) restatement of the reference's ray-rendering
hot path.  TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Every function cites the file:line of ``wagnva/semantic-nerf-for-satellite-data``
it follows (paths relative to the reference root).  The code is written from the
math in SURVEY.md Appendix A, in a functional style over a ``state_dict``-shaped
parameter dictionary (the reference's own parameter names), so the same
parameters can be loaded into the reference ``nn.Module`` for pinning.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "ModelSpec",
    "param_shapes",
    "make_params",
    "linspace01",
    "sample_z",
    "sample_points",
    "posenc",
    "mlp_forward",
    "convert_sigmas",
    "composite",
    "inference",
    "render_rays",
    "synthetic_rays",
    "snerf_loss",
    "satnerf_loss",
    "depth_loss",
    "semantic_loss",
    "car_reg_loss",
    "semantic_uncertainty_loss",
    "batched_inference",
    "xyz_from_depth",
    "denormalize",
    "psnr",
]


# --------------------------------------------------------------------------------------
# model description
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ModelSpec:
    """Architecture of the two models on the path.

    kind = "satnerf":  baseline/models/satnerf.py:101-206 (mapping=False: raw xyz input,
                       baseline/pipelines/satnerf.py:53-59 never overrides the default)
    kind = "semantic": semantic/models/rs_semantic.py:139-258 (positional Mapping always on)
    kind = "nerf":     baseline/models/nerf.py:98-212 (NeRF as baseline/pipelines/nerf.py:26-34 builds it: the
                       constructor defaults mapping=True, siren=False -> positional encoding of xyz (10 freqs) and of the
                       view direction (4 freqs), ReLU activations, outputs [rgb | sigma])
    kind = "snerf":    baseline/models/snerf.py:104-188 (ShadowNeRF as baseline/pipelines/snerf.py:24-32 builds it:
                       SIREN, raw xyz): SatNeRF without the transient-uncertainty head and embedding, 8 outputs
    """

    kind: str = "semantic"
    n_classes: int = 6
    feat: int = 512
    layers: int = 8
    skips: tuple = (4,)
    tau: int = 4
    vocab: int = 50
    n_freq: int = 10
    full_features: bool = False
    siren: bool = True              # activation_function == "siren" (rs_semantic.py:150,158): else every `nl` is a ReLU, the
                                    # first trunk layer loses its w0 = 30 and no SIREN initialiser runs (semantic model only)
    semantic_sigmoid: bool = True  # configs/pipelines/rs_semantic.toml:55
    # head-input variants of the semantic model (rs_semantic.py:186-215; all off in the shipped TOML)
    tj_for_s: bool = False             # use_tj_for_s: the semantic head reads cat(f, t)
    tj_instead_of_beta: bool = False   # use_tj_instead_of_beta: the colour head reads cat(f, t)
    separate_beta_s: bool = False      # use_separate_beta_for_s: a second uncertainty head, output column 9 (rs_semantic.py:228-237)
    separate_tj_s: bool = False        # use_separate_tj_for_semantic: the semantic heads read a second embedding t_s (:300-301,334-335)

    @property
    def k0(self) -> int:
        return 2 * self.n_freq * 3 if self.kind in ("semantic", "nerf") else 3

    @property
    def kdir(self) -> int:   # encoded view direction (NeRF only): mapping_sizes[1] = 4 frequencies (nerf.py:104)
        return 2 * 4 * 3

    @property
    def feat_last(self) -> int:
        return self.feat if self.full_features else self.feat // 2

    @property
    def n_out(self) -> int:
        if self.kind == "nerf":
            return 4
        if self.kind == "snerf":
            return 8
        return 9 + ((self.n_classes + (1 if self.separate_beta_s else 0)) if self.kind == "semantic" else 0)


def param_shapes(spec: ModelSpec) -> Dict[str, tuple]:
    """Parameter names/shapes exactly as the reference ``state_dict`` has them
    (satnerf.py:143-206, rs_semantic.py:176-258)."""
    f, fl, k0 = spec.feat, spec.feat_last, spec.k0
    s: Dict[str, tuple] = {}
    for i in range(spec.layers):
        if i == 0:
            kin = k0
        elif i in spec.skips:
            kin = f + k0
        else:
            kin = f
        s[f"fc_net.{2 * i}.weight"] = (f, kin)
        s[f"fc_net.{2 * i}.bias"] = (f,)
    s["sigma_from_xyz.0.weight"] = (1, f)
    s["sigma_from_xyz.0.bias"] = (1,)
    s["feats_from_xyz.weight"] = (f, f)
    s["feats_from_xyz.bias"] = (f,)
    s["rgb_from_xyzdir.0.weight"] = (fl, f + (spec.kdir if spec.kind == "nerf" else 0) +
                                     (spec.tau if (spec.kind == "semantic" and spec.tj_instead_of_beta) else 0))
    s["rgb_from_xyzdir.0.bias"] = (fl,)
    s["rgb_from_xyzdir.2.weight"] = (3, fl)
    s["rgb_from_xyzdir.2.bias"] = (3,)
    if spec.kind == "nerf":   # nerf.py:140-160: no further heads
        return s
    if spec.kind == "semantic":
        s["semantic_prediction.0.weight"] = (fl, f + (spec.tau if spec.tj_for_s else 0))
        s["semantic_prediction.0.bias"] = (fl,)
        s["semantic_prediction.2.weight"] = (spec.n_classes, fl)
        s["semantic_prediction.2.bias"] = (spec.n_classes,)
    s["sun_v_net.0.weight"] = (fl, f + 3)
    s["sun_v_net.0.bias"] = (fl,)
    for j in (2, 4):
        s[f"sun_v_net.{j}.weight"] = (fl, fl)
        s[f"sun_v_net.{j}.bias"] = (fl,)
    s["sun_v_net.6.weight"] = (1, fl)
    s["sun_v_net.6.bias"] = (1,)
    s["sky_color.0.weight"] = (fl, 3)
    s["sky_color.0.bias"] = (fl,)
    s["sky_color.2.weight"] = (3, fl)
    s["sky_color.2.bias"] = (3,)
    if spec.kind == "snerf":   # snerf.py:161-186: no beta head
        return s
    s["beta_from_xyz.0.weight"] = (fl, f + spec.tau)
    s["beta_from_xyz.0.bias"] = (fl,)
    s["beta_from_xyz.2.weight"] = (1, fl)
    s["beta_from_xyz.2.bias"] = (1,)
    if spec.kind == "semantic" and spec.separate_beta_s:
        s["semantic_beta_from_xyz.0.weight"] = (fl, f + spec.tau)
        s["semantic_beta_from_xyz.0.bias"] = (fl,)
        s["semantic_beta_from_xyz.2.weight"] = (1, fl)
        s["semantic_beta_from_xyz.2.bias"] = (1,)
    return s


def make_params(spec: ModelSpec, seed: int = 0, dtype=torch.float32, trained_like: bool = False):
    """Deterministic parameters drawn with numpy's PCG64 (stable across torch
    versions, so golden vectors regenerate anywhere).  Distributions follow the
    reference initialisers: SIREN init on ``fc_net`` and ``sun_v_net``
    (commons.py:5-18, rs_semantic.py:239-243), PyTorch ``nn.Linear`` default
    (U(+-1/sqrt(fan_in)) for weight and bias) elsewhere, N(0,1) embedding.

    trained_like=True widens the head weights so that composited outputs have
    non-trivial dynamic range (used for bf16 statistical parity).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    params: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(spec).items():
        if name.endswith(".bias"):
            wshape = param_shapes(spec)[name[:-4] + "weight"]
            bound = 1.0 / math.sqrt(wshape[1])
        else:
            fan_in = shape[1]
            # NeRF is built with siren=False (baseline/pipelines/nerf.py:28-32): nn.Linear default init everywhere
            siren_net = (name.startswith("fc_net.") or name.startswith("sun_v_net.")) and spec.kind != "nerf" and spec.siren
            if siren_net:
                first = name in ("fc_net.0.weight", "sun_v_net.0.weight")
                bound = 1.0 / fan_in if first else math.sqrt(6.0 / fan_in)
            else:
                bound = 1.0 / math.sqrt(fan_in)
                if trained_like and name.split(".")[0] in (
                    "rgb_from_xyzdir", "semantic_prediction", "sigma_from_xyz", "beta_from_xyz", "sky_color"):
                    bound *= 4.0
        arr = rng.uniform(-bound, bound, size=shape)
        params[name] = torch.from_numpy(arr).to(dtype)
    emb = torch.from_numpy(rng.standard_normal(size=(spec.vocab, spec.tau))).to(dtype)
    return params, emb


# --------------------------------------------------------------------------------------
# K1: sampling + encoding
# --------------------------------------------------------------------------------------
def linspace01(n: int, dtype=torch.float32, device=None) -> torch.Tensor:
    """torch.linspace(0, 1, n) as the reference calls it
    (framework/components/rendering.py:95).  Kept as a separate function so the
    CUDA kernel's closed form (i*step for the lower half, 1-(n-1-i)*step for the
    upper half) can be pinned against it."""
    return torch.linspace(0, 1, n, dtype=dtype, device=device)   # on the rays' device, as the reference does


def sample_z(rays: torch.Tensor, n_samples: int, u: Optional[torch.Tensor]) -> torch.Tensor:
    """Stratified depths along each ray.  framework/components/rendering.py:95-110
    with ``use_disp=False, perturb=1.0``; ``u`` is the U[0,1) jitter the reference
    draws with ``torch.rand_like`` (u=None -> no perturbation, the perturb=0 branch)."""
    near, far = rays[:, 6:7], rays[:, 7:8]
    t = linspace01(n_samples, rays.dtype, rays.device)
    z = near * (1 - t) + far * t
    if u is not None:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * u
    return z


def sample_points(origins: torch.Tensor, dirs: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """xyz = o + d*z  (framework/components/rendering.py:113-115; the solar-correction
    pass uses dirs = sun_d with the same z: semantic/components/rendering.py:61-63)."""
    return origins.unsqueeze(1) + dirs.unsqueeze(1) * z.unsqueeze(2)


def posenc(x: torch.Tensor, n_freq: int = 10) -> torch.Tensor:
    """[sin(2^k x), cos(2^k x)]_{k<n_freq}, no identity term.  baseline/models/commons.py:54,68-74."""
    out = []
    for k in range(n_freq):
        f = float(2.0 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


# --------------------------------------------------------------------------------------
# K2: the MLP
# --------------------------------------------------------------------------------------
def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def spec_for_case(name: str, kind: str, n_classes: int, feat: int = 512) -> ModelSpec:
    """the architecture a golden case's NAME asks for (oracle/pin_against_reference.py CASES, tests/helpers.py GOLDEN_CASES):
    tokens `tj` (use_tj_for_s + use_tj_instead_of_beta), `bs` (use_separate_beta_for_s), `ts` (use_tj_for_s +
    use_separate_beta_for_s + use_separate_tj_for_semantic), `full` (fc_use_full_features), `tauN` (t_embedding_tau = N),
    `relu` (activation_function = "relu"), `freqN` (mapping_pos_n_freq = N)"""
    tok = name.split("_")
    tj, ts, bs = "tj" in tok, "ts" in tok, "bs" in tok
    tau = next((int(t[3:]) for t in tok if t.startswith("tau") and t[3:].isdigit()), 4)
    n_freq = next((int(t[4:]) for t in tok if t.startswith("freq") and t[4:].isdigit()), 10)
    return ModelSpec(kind=kind, n_classes=n_classes, feat=feat, tau=tau, n_freq=n_freq, full_features="full" in tok, siren="relu" not in tok,
                     tj_for_s=tj or ts,
                     tj_instead_of_beta=tj, separate_beta_s=bs or ts, separate_tj_s=ts)


def make_emb_s(spec: ModelSpec, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    """the second embedding table models["t_s"] of `use_separate_tj_for_semantic` (semantic/pipelines/rs_semantic.py:72-77):
    N(0,1) like nn.Embedding's initialiser, from its own deterministic stream"""
    rng = np.random.Generator(np.random.PCG64(seed + 4242))
    return torch.from_numpy(rng.standard_normal(size=(spec.vocab, spec.tau))).to(dtype)


def mlp_forward(p: Dict[str, torch.Tensor], spec: ModelSpec, xyz: torch.Tensor,
                sun_d: torch.Tensor, t: torch.Tensor, return_hidden: bool = False, t_s: Optional[torch.Tensor] = None):
    """(B,3),(B,3),(B,tau) -> (B, 9[+C]) packed [rgb 0:3 | sigma 3 | sun 4 | sky 5:8 | beta 8 | sem 9:].
    satnerf.py:208-255 / rs_semantic.py:260-340; Siren: commons.py:27-38 (w0=30 on the
    first trunk layer only, satnerf.py:146)."""
    if spec.kind == "nerf":
        return _nerf_forward(p, spec, xyz, sun_d, return_hidden)   # `sun_d` carries the view direction here
    enc = posenc(xyz, spec.n_freq) if spec.kind == "semantic" else xyz
    nl = torch.sin if spec.siren else torch.relu      # `nl` of rs_semantic.py:158 / satnerf.py:127
    h = enc
    hidden = []
    for i in range(spec.layers):
        if i in spec.skips:
            h = torch.cat([enc, h], -1)
        y = _lin(p, f"fc_net.{2 * i}", h)
        h = nl(30.0 * y) if (i == 0 and spec.siren) else nl(y)
        hidden.append(h)
    sigma = F.softplus(_lin(p, "sigma_from_xyz.0", h))
    f = _lin(p, "feats_from_xyz", h)
    f_rgb = torch.cat([f, t], -1) if (spec.kind == "semantic" and spec.tj_instead_of_beta) else f   # rs_semantic.py:287-288
    rgb = torch.sigmoid(_lin(p, "rgb_from_xyzdir.2", nl(_lin(p, "rgb_from_xyzdir.0", f_rgb))))
    rgb = rgb * (1 + 2 * 0.001) - 0.001
    s = torch.cat([f, sun_d], -1)
    s = nl(_lin(p, "sun_v_net.0", s))
    s = nl(_lin(p, "sun_v_net.2", s))
    s = nl(_lin(p, "sun_v_net.4", s))
    sun_v = torch.sigmoid(_lin(p, "sun_v_net.6", s))
    sky = torch.sigmoid(_lin(p, "sky_color.2", torch.relu(_lin(p, "sky_color.0", sun_d))))
    if spec.kind == "snerf":   # snerf.py:226-242: [rgb | sigma | sun_v | sky]
        out = torch.cat([rgb, sigma, sun_v, sky], 1)
        return (out, hidden, f) if return_hidden else out
    beta = F.softplus(_lin(p, "beta_from_xyz.2", nl(_lin(p, "beta_from_xyz.0", torch.cat([f, t], -1)))))
    cols = [rgb, sigma, sun_v, sky, beta]
    # the semantic heads' embedding: t, or the separate t_s (rs_semantic.py:300-301,334-335)
    t_sem = t_s if (spec.kind == "semantic" and spec.separate_tj_s) else t
    if spec.kind == "semantic" and spec.separate_beta_s:   # rs_semantic.py:297-303
        cols.append(F.softplus(_lin(p, "semantic_beta_from_xyz.2",
                                    nl(_lin(p, "semantic_beta_from_xyz.0", torch.cat([f, t_sem], -1))))))
    if spec.kind == "semantic":
        f_sem = torch.cat([f, t_sem], -1) if spec.tj_for_s else f                                      # rs_semantic.py:330-338
        sem = _lin(p, "semantic_prediction.2", nl(_lin(p, "semantic_prediction.0", f_sem)))
        if spec.semantic_sigmoid:
            sem = torch.sigmoid(sem)
        cols.append(sem)
    out = torch.cat(cols, 1)
    if return_hidden:
        return out, hidden, f
    return out


def _nerf_forward(p, spec: ModelSpec, xyz: torch.Tensor, view_d: torch.Tensor, return_hidden: bool = False):
    """NeRF.forward, baseline/models/nerf.py:164-212 (mapping on, ReLU): (B,3),(B,3) -> (B,4) [rgb | sigma]."""
    enc = posenc(xyz, spec.n_freq)
    h = enc
    hidden = []
    for i in range(spec.layers):
        if i in spec.skips:
            h = torch.cat([enc, h], -1)
        h = torch.relu(_lin(p, f"fc_net.{2 * i}", h))
        hidden.append(h)
    sigma = F.softplus(_lin(p, "sigma_from_xyz.0", h))
    f = _lin(p, "feats_from_xyz", h)
    x = torch.cat([f, posenc(view_d, 4)], -1)
    rgb = torch.sigmoid(_lin(p, "rgb_from_xyzdir.2", torch.relu(_lin(p, "rgb_from_xyzdir.0", x))))
    rgb = rgb * (1 + 2 * 0.001) - 0.001
    out = torch.cat([rgb, sigma], 1)
    return (out, hidden, f) if return_hidden else out


# --------------------------------------------------------------------------------------
# K3: compositing
# --------------------------------------------------------------------------------------
def convert_sigmas(sigmas: torch.Tensor, z: torch.Tensor):
    """framework/util/rendering.py:4-34."""
    deltas = z[:, 1:] - z[:, :-1]
    # NB: exactly as the reference builds it (ones_like of a slice of ``deltas``): for S == 1 that
    # slice is empty and every per-sample tensor collapses to (N,0) - the reference does not support
    # S < 2, and neither does the CUDA path (it returns SNB_ERR_UNSUPPORTED).
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[:, :1])], -1)
    alphas = 1 - torch.exp(-deltas * torch.relu(sigmas))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    transparency = torch.cumprod(shifted, -1)[:, :-1]
    weights = alphas * transparency
    depth = torch.sum(weights * z, -1)
    return weights, depth, transparency, alphas


def composite(out: torch.Tensor, z: torch.Tensor, n_classes: int = 0, separate_beta_s: bool = False) -> Dict[str, torch.Tensor]:
    """Tail of ``inference``: satnerf.py:73-96 / rs_semantic.py:81-126.  ``out`` is (N,S,9[+1][+C])."""
    rgbs, sigmas = out[..., :3], out[..., 3]
    weights, depth, transparency, _ = convert_sigmas(sigmas, z)
    if out.shape[-1] == 4:   # NeRF's inference (nerf.py:73-86): plain emission-absorption, no lighting model, no clamp
        return {"rgb": torch.sum(weights.unsqueeze(-1) * rgbs, -2), "depth": depth, "weights": weights,
                "transparency": transparency}
    sun_v, sky = out[..., 4:5], out[..., 5:8]
    irradiance = sun_v + (1 - sun_v) * sky
    rgb = torch.clamp(torch.sum(weights.unsqueeze(-1) * rgbs * irradiance, -2), min=0.0, max=1.0)
    res = {
        "rgb": rgb, "depth": depth, "weights": weights, "transparency": transparency,
        "albedo": rgbs, "sun": sun_v, "sky": sky,
    }
    if out.shape[-1] > 8:   # SatNeRF / semantic (satnerf.py:73-96); S-NeRF's inference returns neither (snerf.py:86-96)
        res["beta"] = out[..., 8:9]
        res["sigmas"] = sigmas
    if n_classes > 0:
        so = 10 if separate_beta_s else 9                      # rs_semantic.py:90-96
        if separate_beta_s:
            res["beta_semantic"] = out[..., 9:10]
        sem = out[..., so:so + n_classes]
        logits = torch.sum(weights.unsqueeze(-1) * sem, -2)
        res["semantic_logits"] = logits
        # rs_semantic.py:131-136: argmax(softmax(x)) == argmax(x)
        res["semantic_label"] = torch.argmax(torch.softmax(logits, dim=-1), dim=-1)
    return res


def inference(p, spec: ModelSpec, xyz: torch.Tensor, z: torch.Tensor, sun_d: torch.Tensor,
              t: torch.Tensor, t_s: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """satnerf.py:8-98 / rs_semantic.py:8-128 (per-ray inputs broadcast over samples,
    model evaluated on all points, then composited)."""
    n, s = z.shape
    out = mlp_forward(p, spec, xyz.reshape(-1, 3), torch.repeat_interleave(sun_d, s, 0),
                      torch.repeat_interleave(t, s, 0) if t is not None else None,
                      t_s=torch.repeat_interleave(t_s, s, 0) if t_s is not None else None)
    return composite(out.view(n, s, -1), z, spec.n_classes if spec.kind == "semantic" else 0,
                     spec.kind == "semantic" and spec.separate_beta_s)


def render_rays(p, emb: torch.Tensor, spec: ModelSpec, rays: torch.Tensor, extras: torch.Tensor,
                n_samples: int, u: Optional[torch.Tensor] = None, z: Optional[torch.Tensor] = None,
                sc_lambda: float = 0.05, emb_s: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """BaseRenderer.render_rays (framework/components/rendering.py:125-157) +
    {SatNeRF,RSSemantic}Rendering._model_rendering (baseline/components/rendering.py:12-67,
    semantic/components/rendering.py:18-80): main pass, optional solar-correction pass on
    o + sun_d*z, ``_coarse`` suffix on every key."""
    if z is None:
        z = sample_z(rays, n_samples, u)
    o, d = rays[:, 0:3], rays[:, 3:6]
    sun_d = extras[:, 0:3]
    ts = extras[:, 3].long()
    t = emb[ts] if spec.kind not in ("snerf", "nerf") else None   # no embedding (baseline/components/rendering.py:70-118)
    t_s = emb_s[ts] if emb_s is not None else None                # models["t_s"](ts), semantic/components/rendering.py:43-45
    if spec.kind == "nerf":   # NeRFRendering (rendering.py:103-118): view direction instead of sun direction, one pass
        res = inference(p, spec, sample_points(o, d, z), z, d, None)
        out = {f"{k}_coarse": v for k, v in res.items()}
        out["_z_vals"] = z
        return out
    res = inference(p, spec, sample_points(o, d, z), z, sun_d, t, t_s)
    if sc_lambda > 0:
        tmp = inference(p, spec, sample_points(o, sun_d, z), z, sun_d, t, t_s)
        res["weights_sc"] = tmp["weights"]
        res["transparency_sc"] = tmp["transparency"]
        res["sun_sc"] = tmp["sun"]
    out = {f"{k}_coarse": v for k, v in res.items()}
    out["_z_vals"] = z  # not a reference key; convenience for tests
    return out


# --------------------------------------------------------------------------------------
# synthetic inputs - shared by tests, smoke and bench
# --------------------------------------------------------------------------------------
def synthetic_rays(n_rays: int, seed: int = 0, n_images: int = 17, dtype=torch.float32):
    """rays (N,8) = [o, d, near, far]; extras (N,4) = [sun_d, ts]
    (layout: framework/components/rays.py:7-64; value ranges: SURVEY 8d)."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    o = np.concatenate([rng.uniform(-1, 1, (n_rays, 2)), rng.uniform(0.15, 0.35, (n_rays, 1))], 1)
    d = np.array([0.10, 0.05, -1.0]) + 0.02 * rng.standard_normal((n_rays, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    near = np.zeros((n_rays, 1))
    far = rng.uniform(0.45, 0.65, (n_rays, 1))
    el = np.deg2rad(rng.uniform(30, 70, n_images))
    az = np.deg2rad(rng.uniform(100, 200, n_images))
    sun = np.stack([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)], 1)
    ts = rng.integers(0, n_images, n_rays)
    rays = np.concatenate([o, d, near, far], 1)
    extras = np.concatenate([sun[ts], ts[:, None].astype(np.float64)], 1)
    return torch.from_numpy(rays).to(dtype), torch.from_numpy(extras).to(dtype)


# --------------------------------------------------------------------------------------
# losses that sit directly on the path's outputs (used for gradient parity)
# --------------------------------------------------------------------------------------
def _solar_correction(res, lambda_sc):
    """baseline/components/loss.py:4-13."""
    sun_sc = res["sun_sc_coarse"].squeeze(-1)
    term2 = torch.sum(torch.square(res["transparency_sc_coarse"].detach() - sun_sc), -1)
    term3 = 1 - torch.sum(res["weights_sc_coarse"].detach() * sun_sc, -1)
    return lambda_sc / 3.0 * torch.mean(term2) + lambda_sc / 3.0 * torch.mean(term3)


def nerf_loss(res, gt_rgb, lambda_sc=0.0):
    """NerfLoss, baseline/components/loss.py:97-110 (coarse network only)."""
    return F.mse_loss(res["rgb_coarse"], gt_rgb)


def snerf_loss(res, gt_rgb, lambda_sc=0.05):
    """SNerfLoss, baseline/components/loss.py:71-94 (the ``loss_without_beta`` of the first epochs)."""
    loss = F.mse_loss(res["rgb_coarse"], gt_rgb)
    if lambda_sc > 0:
        loss = loss + _solar_correction(res, lambda_sc)
    return loss


def satnerf_loss(res, gt_rgb, lambda_sc=0.05, beta_min=0.05):
    """SatNerfLoss + uncertainty_aware_loss, baseline/components/loss.py:16-27,50-68."""
    beta = torch.sum(res["weights_coarse"].unsqueeze(-1) * res["beta_coarse"], -2) + beta_min
    loss = ((res["rgb_coarse"] - gt_rgb) ** 2 / (2 * beta ** 2)).mean()
    loss = loss + (3 + torch.log(beta).mean()) / 2
    if lambda_sc > 0:
        loss = loss + _solar_correction(res, lambda_sc)
    return loss


def depth_loss(res, depths, weights=1.0, lambda_ds=1000.0):
    """DepthLoss, baseline/components/loss.py:30-47."""
    return lambda_ds / 3.0 * torch.mean(weights * (res["depth_coarse"] - depths) ** 2)


def _apply_mask(logits, labels, ignore_mask):
    """`inputs[...][ignore_mask], targets[ignore_mask].squeeze()` of the reference losses (semantic/components/loss.py:52-55)."""
    labels = labels.reshape(-1).long()
    if ignore_mask is not None:
        m = ignore_mask.reshape(-1).bool()
        logits, labels = logits[m], labels[m]
    return logits, labels


def semantic_loss(res, labels, lambda_s=0.04, ignore_index=-100, ignore_mask=None):
    """SemanticLoss, semantic/components/loss.py:35-65 (CE over the composited class scores; `ignore_mask` is the
    dataset's semantic_sparsity_mask, semantic/dataset/semantic_dataset.py:65,87)."""
    logits, labels = _apply_mask(res["semantic_logits_coarse"], labels, ignore_mask)
    return lambda_s * F.cross_entropy(logits, labels, ignore_index=ignore_index)


def car_reg_loss(res, labels, car_label, lambda_c=0.1, ignore_mask=None):
    """SemanticCarRegLoss, semantic/components/loss.py:117-157.  NB: with no selected ray the reference's MSELoss of an
    empty tensor is NaN; so is this."""
    unc = torch.sum(res["weights_coarse"].unsqueeze(-1) * res["beta_coarse"], -2)
    sel_mask = labels.reshape(-1) == car_label
    if ignore_mask is not None:
        sel_mask = torch.logical_and(sel_mask, ignore_mask.reshape(-1).bool())
    sel = unc[sel_mask]
    return lambda_c * F.mse_loss(torch.ones_like(sel), sel)


def semantic_uncertainty_loss(res, labels, lambda_s=0.04, ignore_index=-100, ignore_mask=None, beta_min=0.05,
                              detach_beta=False):
    """SemanticUncertaintyLoss + uncertainty_aware_semantic_loss, semantic/components/loss.py:6-32,68-114: the (scalar)
    cross-entropy mean divided by 2 beta^2 per ray, then averaged; beta is the composited uncertainty - of the separate
    semantic head when the render returned `beta_semantic_coarse`, which also adds a second log-beta term."""
    beta_in = res.get("beta_semantic_coarse", res["beta_coarse"])
    if detach_beta:
        beta_in = beta_in.detach().clone()
    beta = torch.sum(res["weights_coarse"].unsqueeze(-1) * beta_in, -2) + beta_min
    logits, lab = _apply_mask(res["semantic_logits_coarse"], labels, ignore_mask)
    ce = F.cross_entropy(logits, lab, ignore_index=ignore_index)
    loss = lambda_s * (ce / (2 * beta ** 2)).mean()
    if "beta_semantic_coarse" in res:
        loss = loss + lambda_s * (3 + torch.log(beta).mean()) / 2
    return loss


# --------------------------------------------------------------------------------------
# callers of the path: chunked whole-image inference and point-cloud extraction
# --------------------------------------------------------------------------------------
def batched_inference(p, emb, spec: ModelSpec, rays, extras, n_samples: int, chunk: int, u=None, sc_lambda: float = 0.05):
    """eval/utils/util.py:13-42 (and BaseRayPipeline.forward, baseline/pipelines/base_ray_pipeline.py:34-54): render_rays on
    consecutive chunks of `chunk` rays, every key concatenated along dim 0.  `u` (N,S) is the jitter the reference draws chunk
    by chunk with torch.rand_like."""
    res: Dict[str, list] = {}
    for i in range(0, rays.shape[0], chunk):
        r = render_rays(p, emb, spec, rays[i:i + chunk], extras[i:i + chunk], n_samples,
                        u=None if u is None else u[i:i + chunk], sc_lambda=sc_lambda)
        for k, v in r.items():
            res.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in res.items()}


def xyz_from_depth(rays: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """SatNeRFDataset.get_xyz_from_nerf_prediction, baseline/dataset/satnerf_dataset.py:156-171: the ray end points at the
    predicted depth, in float64 and normalised scene coordinates."""
    rays, depth = rays.double(), depth.double()
    return rays[:, 0:3] + rays[:, 3:6] * depth.view(-1, 1)


def denormalize(xyz_n: torch.Tensor, center, scale: float) -> torch.Tensor:
    """StandardNormalization.denormalize, baseline/components/normalization.py:50-58 (`range` = the largest of the three
    scales, :60-79): xyz * range + center."""
    out = xyz_n * scale
    for c in range(3):
        out[:, c] += center[c]
    return out


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """eval/utils/metrics.py:17-18: -10*log10(mse)."""
    return float(-10.0 * torch.log10(torch.mean((a - b) ** 2)))

"""nn.Modules mirroring the reference models (same constructor inputs, same ``forward`` signature,
same ``state_dict`` names/shapes) whose math runs in libsnb's CUDA kernels.

Reference: baseline/models/satnerf.py:101-255 (SatNeRF), semantic/models/rs_semantic.py:139-340
(RSSemanticNeRF).  Parameters live in ONE flat fp32 ``nn.Parameter`` (so the optimiser step, the
bf16 re-pack and the data-parallel all-reduce are single contiguous passes); ``state_dict()`` /
``load_state_dict()`` expose them under the reference's names (``fc_net.0.weight`` ...), so
reference checkpoints load here and ours load into the reference (SURVEY trap #12).
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from . import _lib
from ._lib import HEADS_ALL, MODEL_NERF, MODEL_SATNERF, MODEL_SEMANTIC, MODEL_SNERF, check, ptr, stream


class SnbMLP(torch.nn.Module):
    """Common base: owns the libsnb model handle, the flat parameter and the packed bf16 image."""

    def __init__(self, kind: int, n_classes: int, semantic_sigmoid: bool, tau: int, cfgs=None, variant: int = 0, n_freq: int = 10):
        super().__init__()
        tau = int(tau)
        max_tau = 6 if (variant & _lib.VARIANT_SEPARATE_TJ_S) else 12
        if kind in (MODEL_SATNERF, MODEL_SEMANTIC) and not 1 <= tau <= max_tau:
            raise _lib.SnbError(f"t_embedding_tau = {tau}: libsnb carries the per-ray head inputs in 16 columns [1 | sun_d | t | t_s], "
                                f"so the embedding may be 1..{max_tau} wide here")
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.snb_model_create(C.byref(h), kind, n_classes, 1 if semantic_sigmoid else 0, variant, tau, int(n_freq)),
              "snb_model_create")
        self._h = h
        self.kind = kind
        self.semantic_n_classes = n_classes          # read by the reference's inference(), rs_semantic.py:95
        self.beta_s = 1 if (variant & _lib.VARIANT_SEPARATE_BETA_S) else 0   # column 9 = the separate semantic uncertainty
        # use_separate_tj_for_semantic: the semantic heads read a second embedding t_s; the kernels take [t | t_s] as ONE
        # (.., 2 tau) tensor (aux columns 4..4+tau and 4+tau..4+2tau), the gradient comes back the same way and autograd splits it
        self.sep_ts = bool(variant & _lib.VARIANT_SEPARATE_TJ_S)
        self.relu = kind == MODEL_NERF or bool(variant & _lib.VARIANT_RELU)   # no SIREN: ReLU activations, nn.Linear default init
        self.number_of_outputs = 9 + self.beta_s + n_classes       # satnerf.py:120 / rs_semantic.py:291-311
        self.n_out_kernel = 9 + self.beta_s + n_classes            # columns of the packed tensor the kernels write
        self.t_embedding_dims = tau
        self.enc_ld = 64 if kind in (MODEL_SATNERF, MODEL_SNERF) else 128
        n = lib.snb_model_param_count(h)
        self.table: List[Tuple[str, int, Tuple[int, ...]]] = []
        for i in range(lib.snb_model_num_tensors(h)):
            name, off, r, c = C.c_char_p(), C.c_int64(), C.c_int(), C.c_int()
            check(lib.snb_model_tensor_info(h, i, C.byref(name), C.byref(off), C.byref(r), C.byref(c)), "tensor_info")
            shape = (r.value, c.value) if c.value else (r.value,)
            self.table.append((name.value.decode(), off.value, shape))
        self.flat = torch.nn.Parameter(torch.zeros(n, dtype=torch.float32))
        self.reset_parameters()
        self._packed = None
        self._packed_version = None

    # ---- initialisation: same distributions as the reference -------------------------------------
    def reset_parameters(self):
        """SIREN init on fc_net / sun_v_net (commons.py:5-18, rs_semantic.py:239-243: U(+-sqrt(6/fan_in)),
        first layers U(+-1/fan_in)); nn.Linear default init elsewhere (biases everywhere)."""
        with torch.no_grad():
            views = self.named_tensors()
            for name, t in views.items():
                if name.endswith(".bias"):
                    fan_in = views[name[:-4] + "weight"].shape[1]
                    b = 1.0 / math.sqrt(fan_in)
                    t.uniform_(-b, b)
                    continue
                fan_in = t.shape[1]
                if (name.startswith("fc_net.") or name.startswith("sun_v_net.")) and not self.relu:
                    first = name in ("fc_net.0.weight", "sun_v_net.0.weight")
                    b = 1.0 / fan_in if first else math.sqrt(6.0 / fan_in)
                else:
                    b = 1.0 / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) bound of nn.Linear
                t.uniform_(-b, b)

    def named_tensors(self) -> "OrderedDict[str, torch.Tensor]":
        """Reference-named views into the flat parameter."""
        out = OrderedDict()
        for name, off, shape in self.table:
            n = int(torch.tensor(shape).prod())
            out[name] = self.flat.detach()[off:off + n].view(shape)  # detach(): shares the version counter
        return out

    def named_grads(self) -> Dict[str, torch.Tensor]:
        g = self.flat.grad
        out = {}
        for name, off, shape in self.table:
            n = int(torch.tensor(shape).prod())
            out[name] = None if g is None else g[off:off + n].view(shape)
        return out

    def offset_of(self, name: str) -> int:
        for n_, off, _ in self.table:
            if n_ == name:
                return off
        raise KeyError(name)

    # ---- state_dict under the reference's names ---------------------------------------------------
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name, t in self.named_tensors().items():
            destination[prefix + name] = t if keep_vars else t.detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        views = self.named_tensors()
        for name, t in views.items():
            key = prefix + name
            if key not in state_dict:
                missing_keys.append(key)
                continue
            src = state_dict[key]
            if tuple(src.shape) != tuple(t.shape):
                error_msgs.append(f"size mismatch for {key}: checkpoint {tuple(src.shape)} vs model {tuple(t.shape)}")
                continue
            with torch.no_grad():
                t.copy_(src)
        if strict:
            for key in state_dict.keys():
                if key.startswith(prefix) and key[len(prefix):] not in views:
                    unexpected_keys.append(key)

    # ---- packed bf16 weights ----------------------------------------------------------------------------
    def packed(self) -> torch.Tensor:
        """bf16 weight image for the GEMMs; re-packed whenever the flat parameter was modified."""
        flat = self.flat
        if not flat.is_cuda:
            raise _lib.SnbError("the model must be on a CUDA device (libsnb has no CPU path)")
        key = (flat.data_ptr(), flat._version)
        if self._packed is None or self._packed.device != flat.device or self._packed_version != key:
            lib = _lib.load()
            if self._packed is None or self._packed.device != flat.device:
                nbytes = lib.snb_model_packed_bytes(self._h)
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
            check(lib.snb_model_pack(self._h, ptr(flat.data), ptr(self._packed), stream()), "snb_model_pack")
            self._packed_version = key
        return self._packed

    def mark_dirty(self):
        """call after the parameters were modified behind PyTorch's back (snb_adam_step on raw pointers)."""
        self._packed_version = None

    def sky_params(self):
        v = self.named_tensors()
        return (v["sky_color.0.weight"], v["sky_color.0.bias"], v["sky_color.2.weight"], v["sky_color.2.bias"])

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().snb_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- reference forward signature ----------------------------------------------------------------
    def forward(self, input_xyz, input_dir=None, input_sun_dir=None, input_t=None, input_t_s=None, epoch=None):
        """(B,3),(B,3),(B,tau) -> (B, 9[+C]) packed exactly like the reference's forward
        (satnerf.py:208-255, rs_semantic.py:260-313)."""
        from .autograd import mlp_fp32, mlp_points
        if self.sep_ts:
            if input_t_s is None:
                raise _lib.SnbError("use_separate_tj_for_semantic: forward() needs input_t_s (rs_semantic.py:300-301)")
            input_t = torch.cat([input_t, input_t_s], 1)
        if getattr(self, "precision", "bf16") == "fp32":   # verification mode: fp32 end to end, inference only
            lib = _lib.load()
            xyz = input_xyz.float().contiguous()
            sun, tt = input_sun_dir.float().contiguous(), input_t.detach().float().contiguous()
            P, dev = xyz.shape[0], xyz.device
            # K1's point encoder supplies the per-point sky colour (fp32); its bf16 encodings are not used in this mode
            enc = torch.empty(P, self.enc_ld, dtype=torch.bfloat16, device=dev)
            aux = torch.empty(P, 16, dtype=torch.bfloat16, device=dev)
            sky = torch.empty(P, 3, dtype=torch.float32, device=dev)
            w1, b1, w2, b2 = self.sky_params()
            check(lib.snb_encode_points(ptr(xyz), ptr(sun), ptr(tt), tt.shape[1], ptr(w1), ptr(b1), ptr(w2), ptr(b2),
                                        w1.shape[0], P, self.kind, ptr(enc), ptr(aux), ptr(sky), stream()), "snb_encode_points")
            return mlp_fp32(self, xyz, sun, tt, sky, 0, HEADS_ALL)
        return mlp_points(self, input_xyz, input_sun_dir, input_t)


class SatNeRFB200(SnbMLP):
    """Drop-in for baseline.models.satnerf.SatNeRF (constructor signature: satnerf.py:101-112)."""

    def __init__(self, cfgs=None, layers=8, feat=512, mapping=False, mapping_sizes=(10, 4), skips=(4,), siren=True,
                 t_embedding_dims=4):
        if layers != 8 or feat != 512 or list(skips) != [4] or mapping:
            raise _lib.SnbError("libsnb implements the shipped SatNeRF configuration: 8x512, skip [4], "
                                "raw-xyz input (configs/pipelines/satnerf.toml)")
        # fc_use_full_features (satnerf.py:123-124): the head hidden layers and sky_color are fc_units wide instead of half
        full = cfgs is not None and bool(getattr(cfgs.pipeline, "fc_use_full_features", False))
        super().__init__(MODEL_SATNERF, 0, True, t_embedding_dims, cfgs,
                         (_lib.VARIANT_FULL_FEATURES if full else 0) | (0 if siren else _lib.VARIANT_RELU))   # satnerf.py:127,146
        self.layers, self.skips = layers, list(skips)


class RSSemanticNeRFB200(SnbMLP):
    """Drop-in for semantic.models.rs_semantic.RSSemanticNeRF (constructor: rs_semantic.py:139-141)."""

    def __init__(self, cfgs, dataset_semantic):
        p = cfgs.pipeline
        if p.fc_layers != 8 or p.fc_units != 512 or list(p.fc_skips) != [4] or not 1 <= p.mapping_pos_n_freq <= 10:
            raise _lib.SnbError("libsnb implements the shipped rs_semantic.toml trunk: 8 x 512, skip [4], 1..10 positional frequencies")
        sig = p.semantic_activation_function == "sigmoid"
        # head-input variants: t as an extra input of the semantic head / of the colour head (rs_semantic.py:186-215)
        variant = (_lib.VARIANT_TJ_FOR_S if getattr(p, "use_tj_for_s", False) else 0) | \
                  (_lib.VARIANT_TJ_INSTEAD_OF_BETA if getattr(p, "use_tj_instead_of_beta", False) else 0) | \
                  (_lib.VARIANT_SEPARATE_BETA_S if getattr(p, "use_separate_beta_for_s", False) else 0) | \
                  (_lib.VARIANT_SEPARATE_TJ_S if getattr(p, "use_separate_tj_for_semantic", False) else 0) | \
                  (_lib.VARIANT_FULL_FEATURES if getattr(p, "fc_use_full_features", False) else 0) | \
                  (0 if p.activation_function == "siren" else _lib.VARIANT_RELU)     # rs_semantic.py:147-150,158
        super().__init__(MODEL_SEMANTIC, int(dataset_semantic.semantic_n_classes), sig, p.t_embedding_tau, cfgs, variant,
                         p.mapping_pos_n_freq)
        self.cfg = p
        self.layers, self.skips = p.fc_layers, list(p.fc_skips)


class ShadowNeRFB200(SnbMLP):
    """Drop-in for baseline.models.snerf.ShadowNeRF as the S-NeRF pipeline builds it (baseline/pipelines/snerf.py:24-32:
    8x512 SIREN, raw xyz, skip [4]; constructor signature snerf.py:104-112).

    S-NeRF is SatNeRF without the transient-uncertainty head and its embedding (snerf.py:161-186 vs satnerf.py:143-206):
    same trunk, sigma, feats, albedo, sun-visibility and sky heads, outputs [rgb | sigma | sun_v | sky] (8 columns).  The
    library model kind SNB_MODEL_SNERF holds exactly these tensors (fused head layer = [rgb | sun] blocks); the kernels'
    packed rows keep 9 columns, the ninth is unused."""

    def __init__(self, layers=8, feat=512, mapping=False, mapping_sizes=(10, 4), skips=(4,), siren=True):
        if layers != 8 or feat != 512 or list(skips) != [4] or mapping or not siren:
            raise _lib.SnbError("libsnb implements the shipped S-NeRF configuration: 8x512 SIREN, skip [4], raw-xyz input "
                                "(configs/pipelines/snerf.toml, baseline/pipelines/snerf.py:24-32)")
        super().__init__(MODEL_SNERF, 0, True, 4, None)
        self.variant = "snerf"
        self.number_of_outputs = 8                   # snerf.py:117-119
        self.layers, self.skips = layers, list(skips)

    def forward(self, input_xyz, input_dir=None, input_sun_dir=None, sigma_only=False):
        """(B,3), -, (B,3) -> (B,8) [rgb | sigma | sun_v | sky]  (snerf.py:190-243); sigma_only -> (B,1)."""
        t = torch.zeros(input_xyz.shape[0], self.t_embedding_dims, dtype=torch.float32, device=input_xyz.device)
        out = super().forward(input_xyz, input_sun_dir=input_sun_dir, input_t=t)
        return out[:, 3:4] if sigma_only else out[:, :8]


class NeRFB200(SnbMLP):
    """Drop-in for baseline.models.nerf.NeRF as the NeRF pipeline builds it (baseline/pipelines/nerf.py:26-34: constructor
    defaults mapping=True, siren=False -> positional encoding of xyz (10 frequencies) and of the view direction (4), ReLU
    activations, nn.Linear default initialisation; nerf.py:98-162).  Outputs [rgb | sigma] (4 columns).

    Its kernel plans hold the trunk, sigma, feats and the rgb head only; the packed sun column is pinned to 1, so K3's
    lighting model reduces to NeRF's plain emission-absorption sum (nerf.py:73-86).  NB: K3 clamps the composited colour to [0, 1] (as SatNeRF / S-NeRF do); NeRF's inference does not -
    the two differ only when a composited channel leaves [0, 1], by at most 1e-3 (the sigmoid padding, nerf.py:203)."""

    def __init__(self, layers=8, feat=512, mapping=True, mapping_sizes=(10, 4), skips=(4,), siren=False):
        if layers != 8 or feat != 512 or list(skips) != [4] or not mapping or siren or list(mapping_sizes) != [10, 4]:
            raise _lib.SnbError("libsnb implements the NeRF configuration the pipeline builds: 8x512 ReLU, skip [4], "
                                "positional encoding 10 / 4 (baseline/pipelines/nerf.py:26-34)")
        super().__init__(MODEL_NERF, 0, True, 4, None)
        self.variant = "nerf"
        self.number_of_outputs = 4                   # nerf.py:116
        self.layers, self.skips = layers, list(skips)

    def sky_params(self):
        raise _lib.SnbError("NeRF has no sky_color head")

    def forward(self, input_xyz, input_dir=None, sigma_only=False, epoch=None):
        """(B,3),(B,3) -> (B,4) [rgb | sigma]  (nerf.py:164-212); sigma_only -> (B,1)."""
        from .autograd import mlp_fp32, mlp_points
        if getattr(self, "precision", "bf16") == "fp32":   # verification mode: fp32 end to end, inference only
            out = mlp_fp32(self, input_xyz, posenc_dirs(input_dir), None, None, 0, HEADS_ALL)
        else:
            t = torch.zeros(input_xyz.shape[0], self.t_embedding_dims, dtype=torch.float32, device=input_xyz.device)
            out = mlp_points(self, input_xyz, input_dir, t)
        return out[:, 3:4] if sigma_only else out[:, :4]


def posenc_dirs(dirs: torch.Tensor, n_freq: int = 4) -> torch.Tensor:
    """Mapping(4, 3)(dir) in fp32 for the fp32 verification mode: [sin(2^k d), cos(2^k d)]_k (commons.py:68-74)."""
    d = dirs.float()
    return torch.cat([f(float(2 ** k) * d) for k in range(n_freq) for f in (torch.sin, torch.cos)], -1).contiguous()

"""Shared test helpers (config stub, golden loader)."""
from __future__ import annotations

import os
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_cfgs(spec, n_samples: int, sc_lambda: float):
    """The attribute bag the hot path reads from ``cfgs.pipeline`` (same field names as the reference's
    pydantic config: baseline/pipelines/satnerf.py:115-132, semantic/pipelines/rs_semantic.py:125-175)."""
    pl = types.SimpleNamespace(
        n_samples=n_samples, render_chunk_size=40960, fc_units=spec.feat, fc_layers=spec.layers,
        fc_skips=list(spec.skips), fc_use_full_features=spec.full_features, sc_lambda=sc_lambda,
        t_embedding_tau=spec.tau, t_embedding_vocab=spec.vocab, activation_function="siren" if spec.siren else "relu",
        mapping_pos_n_freq=spec.n_freq, mapping_dir_n_freq=4,
        semantic_activation_function="sigmoid" if spec.semantic_sigmoid else "none",
        use_tj_for_s=spec.tj_for_s, use_tj_instead_of_beta=spec.tj_instead_of_beta, use_beta_for_s=False,
        use_separate_beta_for_s=spec.separate_beta_s, use_separate_tj_for_semantic=spec.separate_tj_s)
    return types.SimpleNamespace(pipeline=pl)


# must match oracle/pin_against_reference.py::CASES (name, kind, C, feat, n_rays, n_samples, sc_lambda, seed)
GOLDEN_CASES = [
    ("sem_c6_s64", "semantic", 6, 512, 24, 64, 0.05, 1),
    ("sem_c5_s8", "semantic", 5, 512, 16, 8, 0.05, 2),
    ("sem_c6_s2", "semantic", 6, 512, 8, 2, 0.0, 3),
    ("sem_c6_s128", "semantic", 6, 512, 8, 128, 0.0, 4),
    ("sat_s64", "satnerf", 0, 512, 24, 64, 0.05, 5),
    ("sat_s8_nosc", "satnerf", 0, 512, 16, 8, 0.0, 6),
    ("sem_c6_s64_trained", "semantic", 6, 512, 24, 64, 0.05, 7),
    ("snerf_s64", "snerf", 0, 512, 24, 64, 0.05, 8),
    ("snerf_s8_nosc", "snerf", 0, 512, 16, 8, 0.0, 9),
    ("nerf_s64", "nerf", 0, 512, 24, 64, 0.0, 10),
    ("nerf_s8", "nerf", 0, 512, 16, 8, 0.0, 11),
    ("sem_c6_s8_tj", "semantic", 6, 512, 16, 8, 0.05, 12),      # use_tj_for_s + use_tj_instead_of_beta
    ("sem_c9_s8_bs", "semantic", 9, 512, 16, 8, 0.05, 13),      # use_separate_beta_for_s
    ("sem_c6_s8_ts", "semantic", 6, 512, 16, 8, 0.05, 14),      # use_tj_for_s + use_separate_beta_for_s + use_separate_tj_for_semantic
    # fc_use_full_features (512-wide head hidden layers and sky_color, satnerf.py:123-124) and other embedding widths
    ("sem_c6_s8_full", "semantic", 6, 512, 16, 8, 0.05, 15),
    ("sat_s8_full", "satnerf", 0, 512, 16, 8, 0.05, 16),
    ("sem_c6_s8_tau8", "semantic", 6, 512, 16, 8, 0.05, 17),
    ("sat_s8_tau2", "satnerf", 0, 512, 16, 8, 0.05, 18),
    ("sem_c6_s8_full_tau6_ts", "semantic", 6, 512, 16, 8, 0.05, 19),   # everything at once
    ("sem_c6_s8_relu", "semantic", 6, 512, 16, 8, 0.05, 20),               # activation_function = "relu"
    ("sat_s8_relu", "satnerf", 0, 512, 16, 8, 0.05, 21),                   # SatNeRF(siren=False)
    ("sem_c6_s8_freq6", "semantic", 6, 512, 16, 8, 0.05, 22),              # mapping_pos_n_freq = 6
]


def golden_inputs(case):
    from oracle import render_oracle as O
    name, kind, C, feat, n, s, sc, seed = case
    spec = O.spec_for_case(name, kind, C, feat)
    params, emb = O.make_params(spec, seed=seed, trained_like=name.endswith("trained"))
    rays, extras = O.synthetic_rays(n, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    u = torch.from_numpy(rng.uniform(0, 1, (n, s))).float()
    gold = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    return spec, params, emb, rays, extras, u, gold

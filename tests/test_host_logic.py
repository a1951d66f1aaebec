"""Host-side mirror of the reference interfaces: parameter names/shapes, state_dict round trips,
initialiser ranges, renderer signature.  CPU only."""
import inspect
import math

import pytest
import torch

from oracle import render_oracle as O
from semnerf_b200 import _lib
from semnerf_b200.model import RSSemanticNeRFB200, SatNeRFB200
from semnerf_b200.renderer import B200Renderer
from tests.helpers import make_cfgs


def build(kind, C=6, name=""):
    spec = O.spec_for_case(name, kind, C)
    cfgs = make_cfgs(spec, 64, 0.05)
    if kind == "semantic":
        return spec, RSSemanticNeRFB200(cfgs, type("D", (), {"semantic_n_classes": C})())
    return spec, SatNeRFB200(cfgs, layers=8, feat=512, skips=[4], t_embedding_dims=spec.tau, siren=spec.siren)


@pytest.mark.parametrize("kind,C,name", [("semantic", 6, ""), ("semantic", 5, ""), ("satnerf", 0, ""),
                                         # fc_use_full_features, other embedding widths, every head variant on top
                                         ("semantic", 6, "full"), ("satnerf", 0, "full"), ("semantic", 6, "tau8"),
                                         ("satnerf", 0, "tau2"), ("semantic", 6, "full_tau6_ts"), ("semantic", 9, "full_bs_tj"),
                                         ("semantic", 6, "relu"), ("satnerf", 0, "relu"), ("semantic", 6, "freq6"),
                                         ("semantic", 5, "freq1_full")])
def test_state_dict_names_and_shapes_match_reference(kind, C, name):
    spec, m = build(kind, C, name)
    want = O.param_shapes(spec)   # pinned against the reference modules by oracle/pin_against_reference.py
    sd = m.state_dict()
    assert list(sd.keys()) == list(want.keys())
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    assert m.semantic_n_classes == C and m.number_of_outputs == 9 + C + (1 if spec.separate_beta_s else 0)


def test_state_dict_round_trip_and_strictness():
    spec, m = build("semantic")
    params, _ = O.make_params(spec, seed=4)
    res = m.load_state_dict(params, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    sd = m.state_dict()
    assert all(torch.equal(sd[k], params[k]) for k in params)
    bad = dict(params)
    bad.pop("fc_net.8.weight")
    bad["bogus.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    res = m.load_state_dict(bad, strict=False)
    assert res.missing_keys == ["fc_net.8.weight"]
    # Lightning-style prefixes (framework/pipelines.py:204-214: model_coarse.*)
    holder = torch.nn.Module()
    holder.model_coarse = m
    assert "model_coarse.fc_net.0.weight" in holder.state_dict()


def test_loading_bumps_the_repack_version():
    spec, m = build("satnerf", 0)
    v0 = m.flat._version
    m.load_state_dict(O.make_params(spec, seed=1)[0])
    assert m.flat._version > v0


def test_initialiser_ranges_follow_the_reference():
    torch.manual_seed(0)
    _, m = build("semantic")
    t = m.named_tensors()
    assert t["fc_net.0.weight"].abs().max() <= 1 / 60 + 1e-7                   # first_layer_sine_init
    assert t["fc_net.8.weight"].abs().max() <= math.sqrt(6 / 572) + 1e-7        # sine_init, skip layer
    assert t["fc_net.8.weight"].abs().max() > 0.9 * math.sqrt(6 / 572)
    assert t["sun_v_net.0.weight"].abs().max() <= 1 / 515 + 1e-7
    assert t["rgb_from_xyzdir.0.weight"].abs().max() <= 1 / math.sqrt(512) + 1e-7
    assert t["beta_from_xyz.0.bias"].abs().max() <= 1 / math.sqrt(516) + 1e-7


def test_unsupported_configurations_fail_loudly():
    spec = O.ModelSpec(kind="semantic")
    for field, value in (("t_embedding_tau", 13), ("fc_units", 256), ("fc_layers", 6), ("mapping_pos_n_freq", 11)):
        cfgs = make_cfgs(spec, 64, 0.05)
        setattr(cfgs.pipeline, field, value)
        with pytest.raises(_lib.SnbError):
            RSSemanticNeRFB200(cfgs, type("D", (), {"semantic_n_classes": 6})())
    cfgs = make_cfgs(O.ModelSpec(kind="semantic", separate_tj_s=True, tj_for_s=True, tau=7), 64, 0.05)
    with pytest.raises(_lib.SnbError):    # two embeddings of 7 do not fit the 16 per-ray columns
        RSSemanticNeRFB200(cfgs, type("D", (), {"semantic_n_classes": 6})())
    with pytest.raises(_lib.SnbError):    # more classes than head rows
        RSSemanticNeRFB200(make_cfgs(spec, 64, 0.05), type("D", (), {"semantic_n_classes": 11})())
    with pytest.raises(_lib.SnbError):
        SatNeRFB200(make_cfgs(O.ModelSpec(kind="satnerf"), 64, 0.05), feat=256)


def test_renderer_signature_matches_reference_baserenderer():
    # framework/components/rendering.py:125-133
    sig = inspect.signature(B200Renderer.render_rays)
    assert list(sig.parameters)[1:] == ["models", "rays", "extras", "epoch", "progress", "render_options"]
    sig = inspect.signature(B200Renderer._model_rendering)
    assert list(sig.parameters)[1:] == ["models", "typ", "cfgs", "rays", "extras", "xyz", "z_vals", "rays_d", "epoch",
                                        "progress", "render_options"]


def test_no_cpu_fallback():
    _, m = build("satnerf", 0)
    with pytest.raises(_lib.SnbError):
        m(torch.zeros(4, 3), input_sun_dir=torch.zeros(4, 3), input_t=torch.zeros(4, 4))


def test_snerf_state_dict_is_the_reference_shadow_nerf():
    """ShadowNeRFB200 (library kind SNB_MODEL_SNERF) holds exactly ShadowNeRF's tensors (snerf.py:104-188): SatNeRF's
    without the uncertainty head."""
    from semnerf_b200.model import ShadowNeRFB200
    spec = O.ModelSpec(kind="snerf", n_classes=0)
    m = ShadowNeRFB200(layers=8, feat=512, skips=[4])
    want = O.param_shapes(spec)
    sd = m.state_dict()
    assert list(sd.keys()) == list(want.keys()) and not any(k.startswith("beta_from_xyz") for k in sd)
    assert all(tuple(sd[k].shape) == tuple(want[k]) for k in want)
    assert m.number_of_outputs == 8 and m.n_out_kernel == 9 and m.enc_ld == 64
    assert m.flat.numel() == sum(int(torch.tensor(s).prod()) for s in want.values())   # no hidden parameters
    params, _ = O.make_params(spec, seed=3)
    res = m.load_state_dict(params, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(_lib.SnbError):
        ShadowNeRFB200(layers=8, feat=256)                                        # only the shipped 8x512 configuration


def test_loss_params_struct_matches_the_header():
    """ctypes mirror of snb_loss_params: same field order and types as include/snb.h (12 x 4 bytes)."""
    import ctypes
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "snb.h")).read()
    body = re.search(r"typedef struct snb_loss_params \{(.*?)\} snb_loss_params;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            typ, names = decl.split(None, 1)
            fields += [(n.strip(), typ) for n in names.split(",")]
    got = [(n, "int" if t is ctypes.c_int else "float") for n, t in _lib.LossParams._fields_]
    assert got == fields and ctypes.sizeof(_lib.LossParams) == 4 * len(fields) == 48


def test_render_loss_signature_and_term_order():
    from semnerf_b200.autograd import LOSS_TERMS
    sig = inspect.signature(B200Renderer.render_loss)
    for name in ("models", "rays", "extras", "rgbs", "semantic", "color", "lambda_s", "ignore_index", "lambda_c", "car_label",
                 "depth", "depth_weights", "lambda_ds", "render_options"):
        assert name in sig.parameters, name
    assert LOSS_TERMS == ("color", "logbeta", "semantic", "car_reg", "sc_term2", "sc_term3", "ds", "semantic_logbeta")


def test_nerf_state_dict_is_the_reference_nerf():
    """NeRFB200 exposes exactly the tensors of the reference's NeRF as its pipeline builds it (nerf.py:98-162: mapping on,
    ReLU): rgb_from_xyzdir.0 takes the 512 features + the 24 encoded view-direction values; no sun / sky / beta heads."""
    from semnerf_b200.model import NeRFB200, posenc_dirs
    spec = O.ModelSpec(kind="nerf", n_classes=0)
    m = NeRFB200(layers=8, feat=512, skips=[4])
    want = O.param_shapes(spec)
    sd = m.state_dict()
    assert list(sd.keys()) == list(want.keys()) and all(tuple(sd[k].shape) == tuple(want[k]) for k in want)
    assert tuple(sd["rgb_from_xyzdir.0.weight"].shape) == (256, 536) and tuple(sd["fc_net.8.weight"].shape) == (512, 572)
    assert m.number_of_outputs == 4 and m.n_out_kernel == 9 and m.enc_ld == 128
    d = torch.nn.functional.normalize(torch.randn(5, 3), dim=1)
    assert torch.allclose(posenc_dirs(d), O.posenc(d, 4))            # the fp32-mode direction encoding = Mapping(4, 3)
    with pytest.raises(_lib.SnbError):
        NeRFB200(siren=True)


def test_chained_launch_deals_every_block_of_both_passes_exactly_once():
    """The two-pass chained launch (k2_chain.cuh ChainPass): every 256-row block of the main pass and of the solar pass is
    carried by exactly one SM pair, a pair works through its main-pass blocks first, and with the library's rotation
    (n_blocks0 mod pairs) no pair carries more than the unavoidable maximum - 4 + 3 instead of 4 + 4 blocks at the
    reference's default batch.  Runs the kernel's own schedule
    code on the host (snb_chain_schedule)."""
    import ctypes as C_
    lib = _lib.load()
    cases = [(256, 256, 74), (2048, 2048, 74), (300, 125, 74), (23, 1, 23), (2, 2, 2), (1, 1, 1), (77, 0, 74), (1000, 999, 74),
             (75, 75, 74), (148, 74, 74), (5, 3, 5)]
    for nb0, nb1, sms_pairs in cases:
        n_pairs = min(max(nb0, nb1), sms_pairs)          # chain_launch: one pair per block at most
        shift = nb0 % min(nb0, sms_pairs)                # ChainPlan::begin_pass
        seen = [set(), set()]
        loads = []
        for pair in range(n_pairs):
            buf = (C_.c_int * 4096)()
            n = lib.snb_chain_schedule(nb0, nb1, shift, n_pairs, pair, buf, 4096)
            assert 0 <= n <= 4096
            seq = [(buf[i] >> 24, buf[i] & 0xFFFFFF) for i in range(n)]
            passes = [p for p, _ in seq]
            assert passes == sorted(passes), (nb0, nb1, pair)         # all of pass 0, then all of pass 1
            for p, b in seq:
                assert b < (nb0, nb1)[p] and b not in seen[p], (nb0, nb1, pair, p, b)
                seen[p].add(b)
            loads.append(n)
        assert seen[0] == set(range(nb0)) and seen[1] == set(range(nb1)), (nb0, nb1)
        if n_pairs == sms_pairs:
            ideal = -(-(nb0 + nb1) // n_pairs)
            assert max(loads) <= ideal + (1 if (nb0 % n_pairs) + (nb1 % n_pairs) > n_pairs else 0), (nb0, nb1, max(loads), ideal)
    # the reference's default batch: 1024 rays x 64 samples = 256 blocks per pass on 74 pairs -> 7 blocks per pair, not 8
    loads = []
    for pair in range(74):
        buf = (C_.c_int * 64)()
        loads.append(lib.snb_chain_schedule(256, 256, 256 % 74, 74, pair, buf, 64))
    assert max(loads) == 7 and min(loads) == 6 and sum(loads) == 512

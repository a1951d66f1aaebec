"""-m gpu: the CUDA kernels, called through the C ABI, against the CPU oracle / plain fp32 matmul
on the same seeded inputs.  Tolerances are written next to each assertion."""
import pytest
import torch

from oracle import render_oracle as O
from semnerf_b200 import _lib
from semnerf_b200._lib import check, ptr, stream
from tests.helpers import make_cfgs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib_or_fail():
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device"
    return _lib.load()


# ------------------------------------------------------------------------------------------------
# K2 building block: the tcgen05 GEMM
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 512, 576), (4096, 256, 256), (8192, 1024, 576)])
def test_gemm_kmajor_epilogues(M, N, K):
    lib = _lib_or_fail()
    torch.manual_seed(0)
    A = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV) * 0.1
    ref = A.float() @ B.float().t()
    out = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_LINEAR, ptr(out), None, N, None, ptr(bias),
                            1.0, 1, stream()), "linear")
    # fp32 accumulation of exact bf16 products: only the final bf16 rounding (2^-9 relative) differs
    assert (out.float() - (ref + bias)).abs().max() <= 2 ** -8 * (ref + bias).abs().max()
    o0 = torch.zeros_like(out)
    sgn = torch.zeros(M, N // 32, device=DEV, dtype=torch.int32)
    check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_SIN, ptr(o0), ptr(sgn), N, None, ptr(bias),
                            3.0, 1, stream()), "sin")
    y = 3.0 * (ref + bias)
    assert (o0.float() - torch.sin(y)).abs().max() <= 2 ** -8 + 1e-4            # bf16 rounding of |sin| <= 1
    # the sign mask of the derivative: [cos(y) < 0], one word per 32 columns, bit k = column 2k, bit 16 + k = column 2k + 1
    # (checked away from the zero crossings)
    pos = torch.tensor([(c >> 1) + 16 * (c & 1) for c in range(32)], device=DEV, dtype=torch.int32)
    bits = ((sgn.view(M, N // 32, 1) >> pos) & 1).reshape(M, N).bool()
    cy = torch.cos(y)
    clear = cy.abs() > 1e-3
    assert torch.equal(bits[clear], (cy < 0)[clear])
    mul = torch.randn(M, N, device=DEV).bfloat16()
    o2 = torch.zeros_like(out)
    check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_MUL, ptr(o2), None, N, ptr(mul), None, 1.0, 1,
                            stream()), "mul")
    want = ref * mul.float()
    assert (o2.float() - want).abs().max() <= 2 ** -8 * want.abs().max()
    # SIREN dgrad epilogue: acc * w0 * cos(y), the derivative rebuilt from h = sin(y) (bf16) and its sign mask
    o3 = torch.zeros_like(out)
    check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, N, K, 0, 0, _lib.EPI_MUL, ptr(o3), ptr(sgn), N, ptr(o0), None, 3.0, 1,
                            stream()), "mul siren")
    want = ref * 3.0 * cy
    err = (o3.float() - want).abs()
    # sqrt(1 - h^2) from a bf16 h: absolute error ~ 2^-9 h^2 / |cos| away from the zero crossings of cos, <= ~0.07 at
    # them (here w0 = 3 spreads the phase uniformly, the worst case): <= 2.5 % rms, i.e. gradient cosine >= 0.9996
    assert err.max() <= 0.1 * 3.0 * ref.abs().max() and err.pow(2).mean().sqrt() <= 0.025 * want.pow(2).mean().sqrt()


def test_gemm_n16_rows():
    lib = _lib_or_fail()
    torch.manual_seed(1)
    M, K = 300, 1792
    A = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    B = (torch.randn(16, K, device=DEV) / K ** 0.5).bfloat16()
    out = torch.zeros(M, 16, device=DEV)
    check(lib.snb_gemm_bf16(ptr(A), K, ptr(B), K, M, 16, K, 0, 0, _lib.EPI_F32ROWS, ptr(out), None, 16, None, None, 1.0,
                            1, stream()), "f32rows")
    assert (out - A.float() @ B.float().t()).abs().max() <= 2e-5   # fp32 accumulation order only


@pytest.mark.parametrize("P,Mf,Nf,splits", [(64, 128, 256, 1), (4096, 512, 512, 4), (5000, 256, 64, 3),
                                            (4096, 512, 16, 2), (70000, 1024, 512, 9)])
def test_gemm_wgrad_splitk_accumulates(P, Mf, Nf, splits):
    """dY^T X with the reduction over samples (MN-major operands), split-K, accumulated in place:
    linearity check - running it twice doubles the result."""
    lib = _lib_or_fail()
    torch.manual_seed(2)
    dY = (torch.randn(P, Mf, device=DEV) * 0.1).bfloat16()
    X = torch.randn(P, Nf, device=DEV).bfloat16()
    ref = dY.float().t() @ X.float()
    G = torch.zeros(Mf, Nf, device=DEV)
    for rep in (1, 2):
        check(lib.snb_gemm_bf16(ptr(dY), Mf, ptr(X), Nf, Mf, Nf, P, 1, 1, _lib.EPI_WGRAD, ptr(G), None, Nf, None, None,
                                1.0, splits, stream()), "wgrad")
        assert (G - rep * ref).abs().max() <= 2e-5 * rep * max(1.0, ref.abs().max().item())


def test_gemm_rejects_bad_arguments():
    lib = _lib_or_fail()
    a = torch.zeros(128, 256, device=DEV, dtype=torch.bfloat16)
    assert lib.snb_gemm_bf16(ptr(a), 256, ptr(a), 256, 128, 256, 256, 0, 0, _lib.EPI_LINEAR, ptr(a), None, 256, None,
                             None, 1.0, 4, stream()) == -1    # split-K only with the accumulate epilogue
    assert lib.snb_gemm_bf16(ptr(a), 256, ptr(a), 256, 128, 128, 256, 0, 0, _lib.EPI_LINEAR, ptr(a), None, 128, None,
                             None, 1.0, 1, stream()) == -2    # bf16 epilogues: N must be a multiple of 256
    assert lib.snb_gemm_bf16(None, 256, ptr(a), 256, 128, 128, 256, 0, 0, 1, ptr(a), None, 128, None, None, 1.0, 1,
                             stream()) == -1


# ------------------------------------------------------------------------------------------------
# K1: sampling + encoding
# ------------------------------------------------------------------------------------------------
def _model(kind, C, seed=3, S=64, sc=0.05, trained_like=False, tj=False, bs=False, ts=False, spec=None):
    """ts: every head variant at once incl. the second embedding (the caller adds models["t_s"] = O.make_emb_s(spec, seed));
    spec: a ready ModelSpec (fc_use_full_features, t_embedding_tau, ...) instead of the flags"""
    from semnerf_b200.model import RSSemanticNeRFB200, SatNeRFB200
    if spec is None:
        spec = O.ModelSpec(kind=kind, n_classes=C, tj_for_s=tj or ts, tj_instead_of_beta=tj, separate_beta_s=bs or ts, separate_tj_s=ts)
    params, emb = O.make_params(spec, seed=seed, trained_like=trained_like)
    cfgs = make_cfgs(spec, S, sc)
    if kind == "snerf":
        from semnerf_b200.model import ShadowNeRFB200
        model = ShadowNeRFB200().to(DEV)
    elif kind == "nerf":
        from semnerf_b200.model import NeRFB200
        model = NeRFB200().to(DEV)
    else:
        model = (RSSemanticNeRFB200(cfgs, type("D", (), {"semantic_n_classes": C})()) if kind == "semantic"
                 else SatNeRFB200(cfgs, t_embedding_dims=spec.tau, siren=spec.siren)).to(DEV)
    model.load_state_dict(params)
    t = torch.nn.Embedding(spec.vocab, spec.tau).to(DEV)
    t.weight.data.copy_(emb)
    return spec, params, emb, cfgs, model, t


@pytest.mark.parametrize("kind", ["semantic", "satnerf"])
@pytest.mark.parametrize("S", [2, 8, 64, 128])
def test_k1_sampling_bit_exact_and_encoding(kind, S):
    from semnerf_b200.autograd import encode_rays
    _lib_or_fail()
    spec, params, emb, cfgs, model, t = _model(kind, 6 if kind == "semantic" else 0, S=S)
    n = 200
    rays, extras = O.synthetic_rays(n, seed=S)
    u = torch.rand(n, S, generator=torch.Generator().manual_seed(S))
    z, enc, enc_sc, aux, sky = encode_rays(model, t.weight, rays.to(DEV), extras.to(DEV), S, u=u.to(DEV), want_sc=True)
    z_ref = O.sample_z(rays, S, u)
    assert torch.equal(z.cpu(), z_ref)                       # bit-exact stratified depths
    k0 = spec.k0
    for e, dirs in ((enc, rays[:, 3:6]), (enc_sc, extras[:, :3])):
        xyz = O.sample_points(rays[:, :3], dirs, z_ref).reshape(-1, 3)
        ref = O.posenc(xyz, 10) if kind == "semantic" else xyz
        e = e.float().cpu()
        if kind == "semantic":    # row = [hi(60) | lo(60) | 0(8)]
            hi, lo, pad = e[:, :k0], e[:, k0:2 * k0], e[:, 2 * k0:]
        else:                     # row = [hi(3) | hi(3) | lo(3) | 0]
            hi, lo, pad = e[:, :k0], e[:, 2 * k0:3 * k0], e[:, 3 * k0:]
            assert torch.equal(e[:, k0:2 * k0], hi)
        assert (hi + lo - ref).abs().max() <= 2 ** -16   # two-term bf16 split: 16 mantissa bits
        assert (pad == 0).all()
    a = aux.float().cpu().view(n, S, 16)
    ref_aux = torch.cat([torch.ones(n, 1), extras[:, :3], emb[extras[:, 3].long()], torch.zeros(n, 8)], 1)
    assert torch.equal(a[:, 0], ref_aux.bfloat16().float()) and torch.equal(a[:, -1], a[:, 0])
    sky_ref = torch.sigmoid(torch.relu(extras[:, :3] @ params["sky_color.0.weight"].t() + params["sky_color.0.bias"])
                            @ params["sky_color.2.weight"].t() + params["sky_color.2.bias"])
    assert (sky.cpu() - sky_ref).abs().max() <= 1e-6


def test_k1_philox_jitter_properties():
    """in-kernel Philox: deterministic per (seed, global ray, sample), inside its stratification bin,
    independent of how rays are sharded (ray_offset)."""
    from semnerf_b200.autograd import encode_rays
    _lib_or_fail()
    spec, params, emb, cfgs, model, t = _model("semantic", 6)
    n, S = 512, 64
    rays, extras = O.synthetic_rays(n, seed=0)
    r, e = rays.to(DEV), extras.to(DEV)
    z1 = encode_rays(model, t.weight, r, e, S, seed=7)[0]
    z2 = encode_rays(model, t.weight, r, e, S, seed=7)[0]
    z3 = encode_rays(model, t.weight, r, e, S, seed=8)[0]
    assert torch.equal(z1, z2) and not torch.equal(z1, z3)
    lo, hi = O.sample_z(rays, S, torch.zeros(n, S)), O.sample_z(rays, S, torch.ones(n, S))
    assert ((z1.cpu() >= lo) & (z1.cpu() <= hi)).all()
    uu = ((z1.cpu() - lo) / (hi - lo)).flatten()
    assert abs(uu.mean().item() - 0.5) < 0.01 and abs(uu.var().item() - 1 / 12) < 0.005
    # sharding invariance: second half rendered alone with ray_offset = n/2
    zb = encode_rays(model, t.weight, r[n // 2:], e[n // 2:], S, seed=7, ray_offset=n // 2)[0]
    assert torch.equal(zb, z1[n // 2:])


# ------------------------------------------------------------------------------------------------
# K3: compositing, forward + backward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,S,C", [(257, 64, 6), (64, 2, 6), (33, 8, 5), (40, 128, 6), (50, 64, 0), (19, 200, 3)])
def test_k3_composite_forward_backward(n, S, C):
    from semnerf_b200.autograd import Composite
    _lib_or_fail()
    torch.manual_seed(0)
    out = torch.rand(n, S, 9 + C)
    out[..., 3] = torch.rand(n, S) * 40 * (torch.rand(n, S) > 0.3)   # sigma incl. exact zeros
    out[0, :, 3] = 0.0      # empty ray
    out[1, :, 3] = 1e4      # alpha -> 1 at the first sample, transmittance underflows
    out[2, :, :3] = 3.0     # clamp active
    rays, _ = O.synthetic_rays(n, seed=S)
    z = O.sample_z(rays, S, torch.rand(n, S))
    o_ref = out.clone().double().requires_grad_(True)
    ref = O.composite(o_ref, z.double(), C)
    o_gpu = out.to(DEV).requires_grad_(True)
    rgb, depth, w, T, sem, label = Composite.apply(o_gpu, z.to(DEV), C)
    tol = 1e-6   # fp32 vs fp64 oracle: reassociated sums / scan order only
    assert (rgb.cpu() - ref["rgb"]).abs().max() <= tol
    assert (depth.cpu() - ref["depth"]).abs().max() <= tol
    assert (w.cpu() - ref["weights"]).abs().max() <= tol
    assert (T.cpu() - ref["transparency"]).abs().max() <= tol
    if C:
        assert (sem.cpu() - ref["semantic_logits"]).abs().max() <= tol
        assert torch.equal(label.cpu(), ref["semantic_label"])
    g = torch.Generator().manual_seed(5)
    gr, gd, gw, gt = (torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, S, generator=g),
                      torch.randn(n, S, generator=g))
    gs = torch.randn(n, max(C, 1), generator=g)[:, :C]
    lr = (ref["rgb"] * gr).sum() + (ref["depth"] * gd).sum() + (ref["weights"] * gw).sum() + (ref["transparency"] * gt).sum()
    lg = (rgb * gr.to(DEV)).sum() + (depth * gd.to(DEV)).sum() + (w * gw.to(DEV)).sum() + (T * gt.to(DEV)).sum()
    if C:
        lr = lr + (ref["semantic_logits"] * gs).sum()
        lg = lg + (sem * gs.to(DEV)).sum()
    lr.backward()
    lg.backward()
    assert (o_gpu.grad.cpu() - o_ref.grad).abs().max() <= 2e-6 * max(1.0, o_ref.grad.abs().max().item())


def test_k3_rejects_degenerate_sample_counts():
    lib = _lib_or_fail()
    x = torch.zeros(64, device=DEV)
    # S = 1: the reference itself collapses to empty per-sample tensors (framework/util/rendering.py:13)
    assert lib.snb_composite_forward(ptr(x), ptr(x), 4, 1, 15, 6, 0, ptr(x), ptr(x), ptr(x), ptr(x), ptr(x), ptr(x),
                                     stream()) == -2
    assert lib.snb_composite_forward(ptr(x), ptr(x), 0, 64, 15, 6, 0, ptr(x), ptr(x), ptr(x), ptr(x), ptr(x), ptr(x),
                                     stream()) == 0     # empty batch is a no-op

"""Launch the trunk-layer GEMM flavours once each at a training-size M (for ncu captures):
forward (sin + sign mask of the derivative), forward (sin only), dgrad (SIREN derivative from h + mask), wgrad (split-K reduce), bias wgrad (N=16)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semnerf_b200 import _lib, build
from semnerf_b200._lib import check, ptr, stream

build.build()
lib = _lib.load()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = "cuda"
torch.manual_seed(0)
A = (torch.randn(P, 512, device=dev) * 0.5).bfloat16()
W = (torch.randn(512, 512, device=dev) / 512 ** 0.5).bfloat16()
bias = torch.zeros(512, device=dev)
o0 = torch.empty_like(A)
sgn = torch.empty(P, 16, dtype=torch.int32, device=dev)
h = torch.sin(torch.randn(P, 512, device=dev)).bfloat16()
mul = torch.randn(P, 512, device=dev).bfloat16()
G = torch.zeros(512, 512, device=dev)
aux = torch.ones(P, 16, device=dev).bfloat16()
Gb = torch.zeros(512, 16, device=dev)
sms = lib.snb_device_sms()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for r in range(reps + 1):
    ev[0].record()
    check(lib.snb_gemm_bf16(ptr(A), 512, ptr(W), 512, P, 512, 512, 0, 0, _lib.EPI_SIN, ptr(o0), ptr(sgn), 512, None, ptr(bias), 1.0, 1, stream()), "fwd2")
    ev[1].record()
    check(lib.snb_gemm_bf16(ptr(A), 512, ptr(W), 512, P, 512, 512, 0, 0, _lib.EPI_SIN, ptr(o0), None, 512, None, ptr(bias), 1.0, 1, stream()), "fwd1")
    ev[2].record()
    check(lib.snb_gemm_bf16(ptr(A), 512, ptr(W), 512, P, 512, 512, 0, 0, _lib.EPI_MUL, ptr(o0), ptr(sgn), 512, ptr(h), None, 1.0, 1, stream()), "dgrad")
    ev[3].record()
    check(lib.snb_gemm_bf16(ptr(A), 512, ptr(mul), 512, 512, 512, P, 1, 1, _lib.EPI_WGRAD, ptr(G), None, 512, None, None, 1.0, sms // 8, stream()), "wgrad")
    ev[4].record()
    check(lib.snb_gemm_bf16(ptr(A), 512, ptr(aux), 16, 512, 16, P, 1, 1, _lib.EPI_WGRAD, ptr(Gb), None, 16, None, None, 1.0, sms // 4, stream()), "bias")
    ev[5].record()
    torch.cuda.synchronize()
names = ["fwd sin + sign mask", "fwd sin (inference)", "dgrad (siren mul)", "wgrad 512x512", "bias wgrad N=16"]
flop = 2.0 * P * 512 * 512
for i, n in enumerate(names):
    ms = ev[i].elapsed_time(ev[i + 1])
    print(f"{n:18s} {ms * 1e3:8.1f} us   {flop / ms / 1e9 if i < 4 else 0:8.1f} TFLOP/s")

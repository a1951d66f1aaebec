"""HBM bandwidth probes with torch ops: write-only (fill), read-only (sum), copy."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 30   # 2 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)


def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


nb = n * 2
print(f"fill (write-only)   {nb / t(lambda: a.fill_(1.0)) / 1e9:8.1f} GB/s")
print(f"zero (write-only)   {nb / t(lambda: a.zero_()) / 1e9:8.1f} GB/s")
print(f"sum  (read-only)    {nb / t(lambda: a.view(torch.int16).sum()) / 1e9:8.1f} GB/s")
print(f"copy (read+write)   {2 * nb / t(lambda: b.copy_(a)) / 1e9:8.1f} GB/s")
f = a.view(torch.float32)
print(f"mul_ (read+write)   {2 * nb / t(lambda: f.mul_(1.0001)) / 1e9:8.1f} GB/s")

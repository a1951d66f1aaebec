"""what the same box delivers on a bare cuBLAS bf16 GEMM (clock, power, TFLOP/s) - context for the power-capped kernels"""
import threading, time, torch, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(5): a @ b
torch.cuda.synchronize()
samples, stop = [], False
def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)); time.sleep(0.002)
th = threading.Thread(target=poll); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 6000
e0.record()
for _ in range(n): a @ b
e1.record(); torch.cuda.synchronize(); stop = True; th.join()
t = e0.elapsed_time(e1) * 1e-3
sm = sorted(x[0] for x in samples[len(samples) // 2:]); pw = sorted(x[1] for x in samples[len(samples) // 2:])
print(f"cuBLAS bf16 8192^3 x {n}: {2 * 8192 ** 3 * n / t / 1e12:.0f} TFLOP/s over {t:.2f} s; SM clock median {sm[len(sm) // 2]} MHz, power median {pw[len(pw) // 2]:.0f} W")

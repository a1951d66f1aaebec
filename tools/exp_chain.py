"""Per-launch GEMM times of one training step (chained passes by kind, wgrads summed), for same-box A/B experiments.
    [SNB_EXP=n] python tools/exp_chain.py [rays] [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semnerf_b200 import _lib, build, synth
from semnerf_b200.trainer import Trainer, default_cfgs
build.build()
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
cfgs = default_cfgs("semantic", n_samples=64, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=0)
tr = Trainer(cfgs, "semantic", 6, device=dev, car_index=4, seed=0)
rays, extras = synth.make_rays(B, seed=0)
rgbs, labels, _ = synth.make_targets(rays, 6, seed=0)
batch = {k: v.to(dev) for k, v in {"rays": rays, "extras": extras, "rgbs": rgbs, "semantic": labels}.items()}
for i in range(3):
    tr.training_step(batch, epoch=3)
torch.cuda.synchronize()
dump = f"/tmp/snb_prof_{os.getpid()}.txt"
os.environ["SNB_PROF_DUMP"] = dump
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    tr.training_step(batch, epoch=3)
e1.record()
torch.cuda.synchronize()
untimed = e0.elapsed_time(e1) / steps
# SM clock / power while the steps run (NVML, sampled every 2 ms from a thread): the box is power-capped, so a variant that
# moves fewer bytes runs at a higher clock - time alone cannot tell a bandwidth limit from an energy limit
import threading, time, pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.005)
th = threading.Thread(target=poll); th.start()
for i in range(int(os.environ.get('SNB_EXP_STEPS', '250'))):
    tr.training_step(batch, epoch=3)
torch.cuda.synchronize()
stop = True; th.join()
sm = sorted(x[0] for x in samples[len(samples) // 2:]); pw = sorted(x[1] for x in samples[len(samples) // 2:])
clk = f"SM clock median {sm[len(sm) // 2]} MHz, power median {pw[len(pw) // 2]:.0f} W ({len(samples)} samples)"
lib.snb_profile_begin(1)
for i in range(steps):
    tr.training_step(batch, epoch=3)
gms, gl, tl, macs = C.c_double(), C.c_int64(), C.c_int64(), C.c_double()
lib.snb_profile_end(C.byref(gms), C.byref(gl), C.byref(tl), C.byref(macs))
rows = [l.split() for l in open(dump) if not l.startswith("#")]
os.unlink(dump)
per = len(rows) // steps
chains, wg = [], 0.0
for r in rows[:per * steps]:
    epi, us = int(r[1]), float(r[7])
    if epi >= 100:
        chains.append(us)
    else:
        wg += us
nch = len(chains) // steps
avg = [sum(chains[i::nch]) / steps for i in range(nch)]
print(f"SNB_EXP={os.environ.get('SNB_EXP', '0')} rays {B}: step {untimed:.3f} ms; chains (us) {[round(a, 1) for a in avg]} sum {sum(avg) / 1e3:.3f} ms; "
      f"other GEMMs {wg / steps / 1e3:.3f} ms; GEMM total {gms.value / steps:.3f} ms; loss {float(tr._loss_out):.4f}; {clk}")

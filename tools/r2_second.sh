#!/bin/bash
O=gpurun_out
T=${1:-r2b}
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/${T}_pytest_gpu.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err; python - <<PY
import json
d=json.load(open("$O/${T}_bench.json"))
for k in ("value","ms_per_step","e2e","gpu_launches","strong_scaling","batch_1024","early_training_step","gpu_eager_baseline","cpu_baseline","cpu_baseline_config1","render"):
    print(k, d.get(k))
r=d["roofline"]; print({k:r[k] for k in ("achieved","frac","frac_of_burst_peak","gemm_ms_per_step","gemm_launches_per_step","gemm_share_of_step","traffic")})
for k,v in (d.get("hbm_kernels") or {}).items():
    print(k, {kk:(round(vv["frac"],3), round(vv["us"],1)) for kk,vv in v.items() if isinstance(vv,dict)} if isinstance(v,dict) else v)
print(d["clocks"])
PY

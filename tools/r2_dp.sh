#!/bin/bash
# N GPUs: the DP equivalence test, then the bench line
O=gpurun_out
T=${1:-r2f}
N=${2:-2}
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s > $O/${T}_pytest_dp.log 2>&1; echo "pytest dp rc=$?"; tail -6 $O/${T}_pytest_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err; echo "bench rc=$?"; tail -2 $O/${T}_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench_${N}gpu.json"))
for k in ("value","ms_per_step","e2e","gpu_launches","strong_scaling"):
    print(k, d.get(k))
r=d["roofline"]; print({k:r[k] for k in ("achieved","frac","gemm_ms_per_step","gemm_share_of_step")})
PY

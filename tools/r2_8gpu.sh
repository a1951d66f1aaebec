#!/bin/bash
O=gpurun_out
T=${1:-r2j}
N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err; echo "bench rc=$?"; tail -2 $O/${T}_bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/bench_configs.py --views 9 > $O/${T}_configs_${N}gpu.jsonl 2> $O/${T}_configs_${N}gpu.err; echo "configs rc=$?"; tail -2 $O/${T}_configs_${N}gpu.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench_${N}gpu.json"))
for k in ("value","ms_per_step","e2e","gpu_launches","strong_scaling","clocks"):
    print(k, d.get(k))
r=d["roofline"]; print({k:r[k] for k in ("achieved","frac","gemm_ms_per_step","gemm_share_of_step")})
for l in open("$O/${T}_configs_${N}gpu.jsonl"):
    l=l.strip()
    if l.startswith("{"):
        j=json.loads(l); print(j["config"][:70], j["metric"], round(j["value"]), j.get("seconds", j.get("ms_per_step")))
PY

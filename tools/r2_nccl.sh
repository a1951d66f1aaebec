#!/bin/bash
# does the overlapped gradient all-reduce slow the weight-gradient GEMMs down (NCCL CTAs on SMs the persistent kernels want)?
# same box: 1 GPU, then N GPUs with NCCL's default CTA count and with NCCL_MAX_CTAS = 2 / 8
O=gpurun_out
T=${1:-r02r}
N=${2:-2}
L=$O/${T}_nccl_ctas.log
: > $L
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-extras > $O/${T}_tmp.json 2> $O/${T}_tmp.err
  python - "$label" $O/${T}_tmp.json >> $L <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print(sys.argv[1], "n_gpus", d["n_gpus"], "ms_per_step %.3f" % d["ms_per_step"], "rays/s %.0f" % d["value"])
PY
}
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras > $O/${T}_tmp.json 2> $O/${T}_tmp.err
python -c "
import json; d=json.load(open('$O/${T}_tmp.json')); print('1 GPU (graph)', 'ms_per_step %.3f' % d['ms_per_step'])" >> $L
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras --no-graph > $O/${T}_tmp.json 2> $O/${T}_tmp.err
python -c "
import json; d=json.load(open('$O/${T}_tmp.json')); print('1 GPU (no graph)', 'ms_per_step %.3f' % d['ms_per_step'])" >> $L
for rep in 1 2; do
run "default" X=1
run "NCCL_MAX_CTAS=2" NCCL_MAX_CTAS=2
run "NCCL_MAX_CTAS=8" NCCL_MAX_CTAS=8
done
cat $L

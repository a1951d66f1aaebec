#!/bin/bash
# same-box A/B of the chained kernel's bottleneck experiments (SNB_EXPERIMENTS build; results deliberately wrong)
export SNB_EXPERIMENTS=1
O=gpurun_out/${1:-r2d}_exp_chain.log
: > $O
python tools/cublas_ref.py >> $O 2>&1
for e in ${2:-0 1 8 0}; do SNB_EXP=$e timeout 300 python tools/exp_chain.py 8192 4 2>&1 | grep SNB_EXP >> $O; done
cat $O

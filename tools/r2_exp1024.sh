#!/bin/bash
# the chained kernel's bottleneck experiments at the reference's default batch (1024 rays: NOT power-capped, so a removed
# stall shows up as time).  SNB_EXPERIMENTS build; results deliberately wrong.
export SNB_EXPERIMENTS=1
O=gpurun_out/${1:-r02l}_exp_chain_b1024.log
: > $O
for e in ${2:-0 1 2 3 4 8 16 0}; do SNB_EXP=$e SNB_EXP_STEPS=50 timeout 300 python tools/exp_chain.py 1024 8 2>&1 | grep SNB_EXP >> $O; done
cat $O

#!/bin/bash
# A/B of one environment knob on the training step:  tools/r2_ab.sh <tag> <ENVVAR> <value A> <value B>
O=gpurun_out
T=${1:-r02k}; V=$2; A=$3; B=$4
timeout 900 python -m pytest tests -m gpu -q -x -k "solar_rows or fused_loss_step or chained_mlp or gemm or wgrad or gradients" > $O/${T}_pytest_sel.log 2>&1; echo "targeted rc=$?"; tail -3 $O/${T}_pytest_sel.log
for rep in 1 2 3; do
for t in $A $B; do
  echo "$V=$t" | tee -a $O/${T}_ab.log
  env $V=$t SNB_EXP_STEPS=50 timeout 300 python tools/exp_chain.py 1024 8 2>&1 | grep SNB_EXP | tee -a $O/${T}_ab.log
done
done
for t in $A $B $A $B; do
  echo "$V=$t" | tee -a $O/${T}_ab.log
  env $V=$t timeout 300 python tools/exp_chain.py 8192 4 2>&1 | grep SNB_EXP | tee -a $O/${T}_ab.log
done

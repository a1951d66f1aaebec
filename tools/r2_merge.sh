#!/bin/bash
# check of the merged main + solar weight gradients: targeted tests, whole GPU suite, step timings at 8192 / 1024 rays
O=gpurun_out
T=${1:-r02h}
timeout 900 python -m pytest tests -m gpu -q -x -k "solar_rows or fused_loss_step or separate_semantic or snerf_training" > $O/${T}_pytest_merge.log 2>&1; echo "targeted rc=$?"; tail -15 $O/${T}_pytest_merge.log
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest_gpu.log
for i in 1 2; do timeout 300 python tools/exp_chain.py 8192 4 2>&1 | grep SNB_EXP | tee -a $O/${T}_exp.log; done
SNB_EXP_STEPS=50 timeout 300 python tools/exp_chain.py 1024 8 2>&1 | grep SNB_EXP | tee -a $O/${T}_exp.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_ncu_launches_step_b1024_merged.csv \
  python tools/prof_step.py 1024 3 > $O/${T}_ncu_b1024.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# launch list (per-kernel durations) of one training step at the reference's default batch (1024 rays)
O=gpurun_out
T=${1:-r02h}
timeout 300 python tools/prof_step.py 1024 6 > $O/${T}_step_b1024.log 2>&1; echo "plain rc=$?"; tail -3 $O/${T}_step_b1024.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_ncu_launches_step_b1024.csv \
  python tools/prof_step.py 1024 3 > $O/${T}_ncu_b1024.log 2>&1; echo "ncu rc=$?"
SNB_EXP_STEPS=50 timeout 300 python tools/exp_chain.py 1024 8 2>&1 | grep SNB_EXP | tee -a $O/${T}_exp.log

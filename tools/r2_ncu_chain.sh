#!/bin/bash
# full ncu capture (source-level sampling) of the two chained launches (forward main + solar, dgrad main + solar) of one training step
O=gpurun_out
T=${1:-r2c}
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:snb_chain --launch-skip 2 --launch-count 2 -f -o $O/${T}_chains \
  python tools/prof_step.py 8192 2 > $O/${T}_ncu_chain.log 2>&1; echo "ncu rc=$?"; tail -3 $O/${T}_ncu_chain.log; ls -la $O/${T}_chains.ncu-rep

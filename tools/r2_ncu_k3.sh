#!/bin/bash
O=gpurun_out
T=${1:-r2h}
i=0
for pat in "k3_composite_kernelILi2ELb0" "k3_composite_kernelILi2ELb1" "k3_loss_kernelILi2"; do
  i=$((i+1))
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:"$pat" --launch-skip 2 --launch-count 1 -f -o $O/${T}_k3_$i python tools/prof_k13.py 40960 > $O/${T}_ncu_k3_$i.log 2>&1; echo "ncu $pat rc=$?"
done
ls -la $O | head

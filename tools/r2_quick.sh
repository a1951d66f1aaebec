#!/bin/bash
# quick check of a kernel change: the GPU test suite, then per-launch GEMM times of the training step
O=gpurun_out
T=${1:-r2q}
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest_gpu.log
for i in 1 2; do timeout 300 python tools/exp_chain.py 8192 4 2>&1 | grep SNB_EXP | tee -a $O/${T}_exp.log; done
SNB_EXP_STEPS=50 timeout 300 python tools/exp_chain.py 1024 8 2>&1 | grep SNB_EXP | tee -a $O/${T}_exp.log

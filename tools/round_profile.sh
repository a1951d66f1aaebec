#!/bin/bash
# One GPU-box pass that produces everything profiles/ keeps for a round: tests, the bench line (both arms), the ncu launch
# list of the same bench command and a per-launch DRAM / tensor-pipe capture of one training step's GEMM kernels.
#   tools/round_profile.sh <tag>      (outputs under gpurun_out/<tag>_*)
T=${1:-rXX}
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest_gpu.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${T}_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-graph > $O/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__cycles_elapsed.avg,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,lts__t_sectors_srcunit_tex.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:"snb_chain|snb_gemm" --launch-skip 18 --launch-count 18 -f -o $O/${T}_step_gemms \
  python tools/prof_step.py 8192 2 > $O/${T}_ncu_step.log 2>&1; echo "ncu step rc=$?"
ls -la $O | head -20

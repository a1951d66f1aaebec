#!/bin/bash
# round 2, first GPU pass: the new tests, then the rest of the suite, then quick bench lines (direct step, graph, 1024 rays)
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -x -q > $O/r2a_pytest_step.log 2>&1; echo "pytest step rc=$?"; tail -15 $O/r2a_pytest_step.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_step.py > $O/r2a_pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -15 $O/r2a_pytest_rest.log
timeout 600 python bench.py --no-cpu > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"; cut -c1-600 $O/r2a_bench.json
timeout 600 python bench.py --no-cpu --graph > $O/r2a_bench_graph.json 2> $O/r2a_bench_graph.err; echo "bench graph rc=$?"; cut -c1-400 $O/r2a_bench_graph.json
timeout 600 python bench.py --no-cpu --batch 1024 --steps 50 > $O/r2a_bench_1024.json 2> $O/r2a_bench_1024.err; echo "bench 1024 rc=$?"; cut -c1-400 $O/r2a_bench_1024.json
timeout 600 python bench.py --no-cpu --batch 1024 --steps 50 --graph > $O/r2a_bench_1024_graph.json 2> $O/r2a_bench_1024_graph.err; echo "bench 1024 graph rc=$?"; cut -c1-400 $O/r2a_bench_1024_graph.json

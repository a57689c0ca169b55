#!/bin/bash
# GPU list builder bring-up: memory checker on a small case, its parity tests, the rest of the GPU suite, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 400 compute-sanitizer --error-exitcode 7 python -m pytest tests/test_tree_gpu.py -x -q -k "lists_equal and (3000 or 7-64 or 9-4)" ) > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
( time timeout 600 python -m pytest tests/test_tree_gpu.py -q ) > gpurun_out/pytest_tree.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_tree.log
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_tree_gpu.py ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -5 gpurun_out/sanitizer.log; tail -15 gpurun_out/pytest_tree.log; tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
tail -c 600 gpurun_out/bench_n1.err; cut -c1-3000 gpurun_out/bench_n1.json

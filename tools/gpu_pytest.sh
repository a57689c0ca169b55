#!/bin/bash
# GPU suite only (optionally: -k expression as $1)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ${1:+-k "$1"} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "rel err|passed|failed|Error|rc=" gpurun_out/pytest_gpu.log | tail -15

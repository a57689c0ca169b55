#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_integration_gpu.py -m gpu -x -q ) > gpurun_out/r2h_integration.log 2>&1
tail -15 gpurun_out/r2h_integration.log
bash tools/gpu_bench_n1.sh 2>&1 | grep -E "^value|^e2e |^parity|rc=|FAILED"
BENCH_ARGS="--n 100000 --no-stage-baseline --cpu-seconds 5" bash tools/gpu_bench_n1.sh > gpurun_out/r2h_n1e5.log 2>&1; grep -E "^value|^e2e |^parity|rc=|FAILED|L2" gpurun_out/r2h_n1e5.log; cp gpurun_out/r2_bench_n1.json gpurun_out/r2_bench_n1e5.json

#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2m_pytest.log
bash tools/gpu_bench_n1.sh

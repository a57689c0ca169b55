#!/bin/bash
mkdir -p gpurun_out
( GPLUM_B200_EPSP_SPLIT=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_soft_corr_gpu.py tests/test_tree_gpu.py -x -q ) > gpurun_out/pytest_split_forced.log 2>&1
echo "forced rc=$?" >> gpurun_out/pytest_split_forced.log
( timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_split_auto.log 2>&1
echo "auto rc=$?" >> gpurun_out/pytest_split_auto.log
GPLUM_B200_EPSP_SPLIT=0 timeout 120 python tools/shard_probe.py 1 2 4 8 16 > gpurun_out/shard_probe_off.log 2>&1
GPLUM_B200_EPSP_SPLIT=-1 timeout 120 python tools/shard_probe.py 1 2 4 8 16 > gpurun_out/shard_probe_auto.log 2>&1
tail -4 gpurun_out/pytest_split_forced.log; tail -4 gpurun_out/pytest_split_auto.log; cat gpurun_out/shard_probe_off.log gpurun_out/shard_probe_auto.log

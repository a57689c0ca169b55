#!/bin/bash
mkdir -p gpurun_out
( GPLUM_B200_BULK=0 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_tree_gpu.py -m gpu -x -q ) > gpurun_out/r2i_pytest_cpasync.log 2>&1
tail -2 gpurun_out/r2i_pytest_cpasync.log
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2i_pytest.log 2>&1
tail -2 gpurun_out/r2i_pytest.log
for r in 1 2; do
timeout 300 python tools/shard_probe.py 1 8 > gpurun_out/r2i_bulk.log 2>&1
GPLUM_B200_BULK=0 timeout 300 python tools/shard_probe.py 1 8 > gpurun_out/r2i_ldgsts.log 2>&1
echo "== bulk"; cat gpurun_out/r2i_bulk.log; echo "== cp.async"; cat gpurun_out/r2i_ldgsts.log
done

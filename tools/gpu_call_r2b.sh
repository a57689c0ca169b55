#!/bin/bash
# round 2, call B: j-split work lists -- GPU suite, shard probe over the split factor
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
for m in 0 2; do
  GPLUM_B200_SPLIT_M=$m timeout 300 python tools/shard_probe.py 1 2 4 8 16 > gpurun_out/r2d_probe_m$m.log 2>&1
done
tail -6 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_probe_m*.log

#!/bin/bash
# round 2, call A: GPU suite on the deferred candidate re-test; shard probe of the default and the UNROLL=8 build
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python tools/shard_probe.py 1 8 > gpurun_out/r2a_probe_u4.log 2>&1
GPLUM_B200_LIB=$PWD/build_variants/libgplum_b200_u8.so timeout 300 python tools/shard_probe.py 1 8 > gpurun_out/r2a_probe_u8.log 2>&1
tail -8 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_probe_u4.log gpurun_out/r2a_probe_u8.log

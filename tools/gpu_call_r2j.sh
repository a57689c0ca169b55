#!/bin/bash
# Newton step on/off: GPU suite with the shipped (off) build, accuracy of both against FP64 sums, sub-wave probes of
# register / unroll variants
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
timeout 300 python tools/newton_probe.py run off 2>&1 | tail -1
GPLUM_B200_LIB=build_variants/libgplum_b200_newton.so timeout 300 python tools/newton_probe.py run on 2>&1 | tail -1
timeout 600 python tools/newton_probe.py compare off on > gpurun_out/r2_newton_step.txt 2>&1; cat gpurun_out/r2_newton_step.txt
for v in "" _b5u4 _b4u4 _b4u8; do
  echo "variant ${v:-shipped}"
  if [ -z "$v" ]; then timeout 300 python tools/shard_probe.py 1 4 8 16 2>&1 | tail -4
  else GPLUM_B200_LIB=build_variants/libgplum_b200$v.so timeout 300 python tools/shard_probe.py 1 4 8 16 2>&1 | tail -4; fi
done

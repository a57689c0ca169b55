#!/bin/bash
mkdir -p gpurun_out
GPLUM_B200_SPLIT_M=0 timeout 300 python tools/shard_probe.py 1 4 8 16 > gpurun_out/r2f_m0.log 2>&1
timeout 300 python tools/shard_probe.py 4 8 16 > gpurun_out/r2f_c1200.log 2>&1
for c in 600 2400 4000; do
  GPLUM_B200_LIB=$PWD/build_variants/libgplum_b200_c$c.so timeout 300 python tools/shard_probe.py 4 8 16 > gpurun_out/r2f_c$c.log 2>&1
done
for f in m0 c600 c1200 c2400 c4000; do echo "== $f"; cat gpurun_out/r2f_$f.log; done

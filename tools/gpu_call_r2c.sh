#!/bin/bash
# ncu of the 1/8-shard force pass with and without j-split
mkdir -p gpurun_out
for m in 0 2; do
  GPLUM_B200_SPLIT_M=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_pass -s 5 -c 1 -f -o gpurun_out/r2c_shard8_m$m \
     python tools/shard_probe.py 8 > gpurun_out/r2c_ncu_m$m.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2c_shard8_m$m.ncu-rep > gpurun_out/r2c_shard8_m$m.txt 2>&1
done
tail -3 gpurun_out/r2c_ncu_m2.log

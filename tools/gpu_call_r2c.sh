#!/bin/bash
# ncu of the 1/8-shard force pass with segments
mkdir -p gpurun_out
for m in 2; do
  GPLUM_B200_SPLIT_M=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_pass -s 5 -c 1 -f -o gpurun_out/r2e_shard8_m$m \
     python tools/shard_probe.py 8 > gpurun_out/r2e_ncu_m$m.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2e_shard8_m$m.ncu-rep > gpurun_out/r2e_shard8_m$m.txt 2>&1
  ncu -i gpurun_out/r2e_shard8_m$m.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if any(k in a for k in ['sm__cycles_active.','sm__cycles_elapsed.avg ','smsp__inst_executed.avg','smsp__inst_executed.max','smsp__inst_executed.min','smsp__cycles_active.','gr__ctas_launched','smsp__issue_active.']): print(a,b,c)
" >> gpurun_out/r2e_shard8_m$m.txt
done
cat gpurun_out/r2e_shard8_m2.txt | head -60

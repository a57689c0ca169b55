#!/bin/bash
mkdir -p gpurun_out
GPLUM_B200_SNAKE=0 timeout 300 python tools/shard_probe.py 1 2 4 8 16 32 > gpurun_out/r2g_plain.log 2>&1
timeout 300 python tools/shard_probe.py 1 2 4 8 16 32 > gpurun_out/r2g_snake.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2g_pytest.log 2>&1
for f in plain snake; do echo "== $f"; cat gpurun_out/r2g_$f.log; done; tail -3 gpurun_out/r2g_pytest.log

#!/bin/bash
# placement beyond one wave: GPLUM_B200_PLACE = waves up to which passes are placed
for pl in 0 1 2 3 8; do
  GPLUM_B200_PLACE=$pl timeout 300 python tools/shard_probe.py 1 2 3 4 6 8 2>&1 | grep walks | sed "s/^/PLACE=$pl /"
done

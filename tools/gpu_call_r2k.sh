#!/bin/bash
# placed sub-wave passes: traces with and without placement, then the GPU suite
mkdir -p gpurun_out
for pl in 0 1; do
  echo "GPLUM_B200_PLACE=$pl"
  GPLUM_B200_PLACE=$pl timeout 300 python tools/trace_probe.py 8 6 16 12 2>&1 | grep -v "SM of the first"
  for k in 8 16; do mv gpurun_out/trace_k$k.npy gpurun_out/trace_k${k}_place$pl.npy; done
done
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log

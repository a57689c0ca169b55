#!/bin/bash
# 8-GPU box: the placed + fused-pack peer test at 8 ranks, bench at N = 8 with and without the fused pack
mkdir -p gpurun_out
( GPLUM_TEST_WORLD=8 timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "placed_pass" ) > gpurun_out/r2_pytest_multi_w8_fused.log 2>&1
tail -3 gpurun_out/r2_pytest_multi_w8_fused.log
for fuse in 1 0; do
  GPLUM_B200_FUSE_PACK=$fuse timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$fuse bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_fuse$fuse.json 2> gpurun_out/r2_bench_n8_fuse$fuse.err
  python - $fuse <<'PY'
import json,sys
f="gpurun_out/r2_bench_n8_fuse%s.json"%sys.argv[1]
try:
    j=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value %.4g ms/step %.4f parity %s e2e %.3f launches %s" % (j["value"], j["ms_per_step"], j["parity_check"]["ok"], j["e2e"]["ms_per_step"], j["gpu_launches"]))
    for p in j.get("phases_all",[])[:3]: print("   ",p)
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done

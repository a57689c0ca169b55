#!/bin/bash
# 8-GPU box: the 2-rank NCCL parity test, then the bench at N = 8, 4, 2
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2_pytest_multi.log 2>&1
tail -3 gpurun_out/r2_pytest_multi.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2_bench_n$n.json") if l.startswith("{")][-1])
    print($n, "value %.4g ms/step %.4f e2e ms %.3f"%(j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"]))
    for p in j.get("phases_all",[])[:3]: print("   ",p)
except Exception as e:
    print($n, "FAILED", e); print(open("gpurun_out/r2_bench_n$n.err").read()[-1500:])
PY
done

#!/bin/bash
# 8-GPU box: the NCCL parity tests at 8 ranks, then the bench at N = 8, 4, 2 and configs[4] (N = 1e7) on 8 GPUs
# (its 1-GPU points run in tools/gpu_call_final.sh)
mkdir -p gpurun_out
( GPLUM_TEST_WORLD=8 timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2_pytest_multi_w8.log 2>&1
tail -3 gpurun_out/r2_pytest_multi_w8.log
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    j=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "N=%d value %.4g ms/step %.4f e2e ms %.3f parity %s soft_step %s" % (j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["parity_check"]["ok"], (j.get("soft_step") or {}).get("ms_per_step")))
    for p in j.get("phases_all",[])[:2]: print("   ",p)
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  show gpurun_out/r2_bench_n$n.json
done
# configs[4]: wide disk N = 1e7, 0.5-10 AU: strong scaling point at 8 GPUs, weak scaling points 1.25e6 per GPU at 1 and 8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --particles 10000000 --a-in 0.5 --a-out 10 > gpurun_out/r2_bench_cfg4_n8.json 2> gpurun_out/r2_bench_cfg4_n8.err
show gpurun_out/r2_bench_cfg4_n8.json

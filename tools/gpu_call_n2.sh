#!/bin/bash
# 2-GPU box: NCCL parity tests (force pass in three exchange modes, multi-rank soft step), bench at N = 2
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2_pytest_multi.log 2>&1
tail -12 gpurun_out/r2_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2_bench_n2.json") if l.startswith("{")][-1])
    print("value %.4g ms/step %.4f e2e ms %.3f"%(j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"]))
    print("e2e", j["e2e"]); print("e2e_multiwalk", j.get("e2e_multiwalk"))
    print("parity", j["parity_check"]); print("soft_step", j.get("soft_step")); print(j["config"]["l2"])
    for p in j.get("phases_all", []): print("   ", p)
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2_bench_n2.err").read()[-2500:])
PY

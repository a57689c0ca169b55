#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --particles 10000000 --a-in 0.5 --a-out 10 > gpurun_out/r2_bench_cfg4_n2.json 2> gpurun_out/r2_bench_cfg4_n2.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2_bench_cfg4_n2.json") if l.startswith("{")][-1])
    print(j["value"], j["ms_per_step"], j["parity_check"]["ok"]); print(j["e2e"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2_bench_cfg4_n2.err").read()[-2500:])
PY

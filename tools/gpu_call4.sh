#!/bin/bash
# full GPU suite + smoke + bench (both arms)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
tail -c 600 gpurun_out/bench_n1.err; python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
print(json.dumps(j["soft_step"])); print(j["value"], j["e2e"], j["roofline"]["frac"], j["gpu_launches"])
print(open("gpurun_out/bench_ref.json").read()[:300])
PY

#!/bin/bash
mkdir -p gpurun_out
( timeout 300 compute-sanitizer --error-exitcode 7 python -m pytest tests/test_iso_step_gpu.py -x -q -k "fixture or 3000" ) > gpurun_out/sanitizer_iso.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitizer_iso.log
( time timeout 600 python -m pytest tests/test_iso_step_gpu.py -q ) > gpurun_out/pytest_iso.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_iso.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -4 gpurun_out/sanitizer_iso.log; tail -30 gpurun_out/pytest_iso.log
tail -c 800 gpurun_out/bench_n1.err; python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
print(json.dumps(j["soft_step"])); print(j["value"], j["e2e"]["ms_per_step"], j["roofline"]["frac"], j["gpu_launches"])
PY

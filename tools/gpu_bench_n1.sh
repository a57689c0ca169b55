#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ${BENCH_ARGS} ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "rc=$?"; tail -c 1500 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2_bench_n1.json") if l.startswith("{")][-1])
    print("value %.4g ms %.4f frac %.3f kernel_ms %.4f"%(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["kernel_ms"]))
    print("e2e", j["e2e"]); print("e2e_multiwalk", j.get("e2e_multiwalk")); print("parity", j["parity_check"])
    print("soft_step", {k:v for k,v in j["soft_step"].items() if k!="resident"}); print("resident", j["soft_step"]["resident"])
    print("cpu", j.get("cpu_baseline")); print("stage", j.get("cpu_baseline_stage")); print(j["config"]["l2"], j["clocks"])
except Exception as e: print("FAILED", e)
PY

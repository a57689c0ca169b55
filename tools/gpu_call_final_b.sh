#!/bin/bash
# trimmed round-end verification on one GPU (the configs[4] points and the rows f1-f3 capture of tools/gpu_call_final.sh
# are unaffected by the last commits): GPU suite, smoke, both bench arms, ncu launch list, ncu --set full of the force pass
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_pass -s 4 -c 1 -f -o gpurun_out/r2_force_pass_full \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_force_pass_full.ncu-rep > gpurun_out/r2_force_pass_ncu_summary.txt 2>&1
python tools/ncu_traffic.py gpurun_out/r2_force_pass_full.ncu-rep 1000000 512 0.9 1.1 > gpurun_out/r2_traffic.log 2>&1
cp profiles/r2_force_pass_traffic.json gpurun_out/ 2>/dev/null
tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2_bench_n1.json") if l.startswith("{")][-1])
print(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["kernel_ms"], j["gpu_launches"]); print(j["e2e"]["ms_per_step"], j["soft_step"]["ms_per_step"], j["parity_check"]["ok"])
print(open("gpurun_out/r2_bench_ref.json").read()[:200])
PY
head -3 gpurun_out/r2_force_pass_ncu_summary.txt; cat gpurun_out/r2_traffic.log

#!/bin/bash
# round-end verification on one GPU: full GPU suite, smoke, both bench arms, ncu launch list of bench.py,
# ncu --set full of the force pass (from bench.py) and of the rows f1-f3 kernels (tools/profile_rows.py)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tree_coop|walk_kernel|gather_kernel|key_kernel|bbox_kernel|corr_|drift_kernel|kick_kernel|pull_kernel|push_kernel|flags_kernel|item_|totals_|tie_fix|unsort|peer_pack" -s 20 -c 26 -f -o gpurun_out/r2_rows_full \
   python tools/profile_rows.py > gpurun_out/ncu_rows.log 2>&1
# configs[4] (wide disk, 0.5-10 AU) on one GPU: N = 1e7 (the strong-scaling base of the 8-GPU point) and N = 1.25e6 (its weak-scaling base)
timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 --n 10000000 --a-in 0.5 --a-out 10 --no-cpu-baseline --no-stage-baseline > gpurun_out/r2_bench_cfg4_n1.json 2> gpurun_out/r2_bench_cfg4_n1.err
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --n 1250000 --a-in 0.5 --a-out 10 --no-cpu-baseline --no-stage-baseline > gpurun_out/r2_bench_cfg4_weak_n1.json 2> gpurun_out/r2_bench_cfg4_weak_n1.err
for f in r2_bench_cfg4_n1 r2_bench_cfg4_weak_n1; do python - $f <<'PY'
import json,sys
f="gpurun_out/%s.json"%sys.argv[1]
try:
    j=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["parity_check"]["ok"], j["roofline"]["frac"])
except Exception as e:
    print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
done
tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
tail -c 600 gpurun_out/r2_bench_n1.err; python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2_bench_n1.json") if l.startswith("{")][-1])
print(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["kernel_ms"], j["gpu_launches"]); print(j["e2e"]); print(j["parity_check"]["ok"], j.get("cpu_baseline"), j.get("cpu_baseline_stage"))
print(open("gpurun_out/r2_bench_ref.json").read()[:300])
PY
head -12 gpurun_out/r2_force_pass_ncu_summary.txt; cat gpurun_out/r2_traffic.log

#!/bin/bash
# cooperative tree kernel: memory checker on small cases, tree parity tests, bench, ncu launch list + captures
mkdir -p gpurun_out
( timeout 400 compute-sanitizer --error-exitcode 7 python -m pytest tests/test_tree_gpu.py -x -q -k "lists_equal and (3000 or 7-64 or 9-4)" ) > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
( time timeout 600 python -m pytest tests/test_tree_gpu.py -q ) > gpurun_out/pytest_tree.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_tree.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tree_coop|walk_kernel|gather_kernel|key_kernel|moment|split" -c 8 -f -o gpurun_out/tree_full \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tree.log 2>&1
tail -4 gpurun_out/sanitizer.log; tail -8 gpurun_out/pytest_tree.log
tail -c 600 gpurun_out/bench_n1.err; python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
print(json.dumps(j["soft_step"])); print(j["value"], j["e2e"]["ms_per_step"], j["roofline"]["frac"])
PY

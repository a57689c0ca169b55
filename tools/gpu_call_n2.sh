#!/bin/bash
# 2-GPU verification: NCCL/peer parity test + bench at N=2 (peer)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > gpurun_out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
for ex in peer; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex > gpurun_out/bench_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err
done
tail -5 gpurun_out/pytest_multi.log
for ex in peer; do tail -c 300 gpurun_out/bench_n2_$ex.err; cut -c1-400 gpurun_out/bench_n2_$ex.json; echo; done

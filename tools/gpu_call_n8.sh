#!/bin/bash
# 8-GPU box: strong-scaling bench at N=8 and N=4 (peer exchange)
mkdir -p gpurun_out
for n in 8 4; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n${n}_peer.json 2> gpurun_out/bench_n${n}_peer.err
done
for f in n8_peer n4_peer; do tail -c 300 gpurun_out/bench_$f.err; cut -c1-300 gpurun_out/bench_$f.json; echo; done

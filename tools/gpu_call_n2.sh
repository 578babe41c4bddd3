#!/bin/bash
# two ranks on one box: the driver's launch line for N = 2 (frame-parallel, one NCCL broadcast of the identity state)
set -u
out=gpurun_out/r2n2
mkdir -p "$out"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > "$out/bench_n2.json" 2> "$out/bench_n2.err"
echo "rc=$?"; tail -c 1500 "$out/bench_n2.json"; tail -c 600 "$out/bench_n2.err"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > "$out/bench_ref_n2.json" 2> "$out/bench_ref_n2.err"
echo "ref rc=$?"; tail -c 600 "$out/bench_ref_n2.json"

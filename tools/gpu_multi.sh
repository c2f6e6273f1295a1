#!/bin/bash
# Multi-GPU lines (one box, 8 GPUs): weak scaling at 8, and BASELINE configs[2] (64 documents per step, strong scaling) at 8 / 4 / 2 / 1.
mkdir -p gpurun_out
: > gpurun_out/r2_bench_multi.jsonl
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { echo "== $*"; timeout 600 "$@" 2>gpurun_out/last_err.txt | tail -1 >> gpurun_out/r2_bench_multi.jsonl; echo "rc=$?"; tail -c 400 gpurun_out/r2_bench_multi.jsonl | cut -c1-260; }
run $T --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline
run $T --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --total-docs 64 --no-extras --no-cpu-baseline
run $T --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 --total-docs 64 --no-extras --no-cpu-baseline
run $T --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --total-docs 64 --no-extras --no-cpu-baseline
run python bench.py --gpus 1 --steps 5 --warmup 3 --total-docs 64 --no-extras --no-cpu-baseline

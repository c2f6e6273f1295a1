#!/bin/bash
mkdir -p gpurun_out
echo "== bench full"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.txt 2>&1; echo "rc=$?"; tail -c 7000 gpurun_out/bench_full.txt
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.txt 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_ref.txt

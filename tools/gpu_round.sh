#!/bin/bash
mkdir -p gpurun_out
echo "== pytest attention+gemm"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "attention or gemm or stages" > gpurun_out/pytest_attn.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_attn.txt
echo "== attn bench"; timeout 200 python tools/attn_bench.py 2>&1 | tail -4
echo "== trace"; DVD_LIB=dvd_b200/libdvd_b200_trace.so timeout 200 python tools/attn_bench.py --d 256 --trace 2>&1 | tail -10
echo "== gemm graph pair"; timeout 300 python tools/gemm_bench.py --graph --bf16x3 > gpurun_out/gemm_graph_pair.txt 2>&1; echo "rc=$?"; cat gpurun_out/gemm_graph_pair.txt | tail -10
echo "== gemm graph pair bf16"; timeout 300 python tools/gemm_bench.py --graph --bf16 2>&1 | tail -10
echo "== bench x3"; timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_x3.txt 2>&1; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_x3.txt').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline']['achieved'],d['kernel_share'])"

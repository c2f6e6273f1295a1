#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gemm"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "gemm or stages or sampling_bf16x3 or batch_equals" > gpurun_out/pytest_gemm.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gemm.txt
echo "== gemm graph pair"; timeout 300 python tools/gemm_bench.py --graph --bf16x3 > gpurun_out/gemm_graph_pair.txt 2>&1; echo "rc=$?"; cat gpurun_out/gemm_graph_pair.txt | tail -10
for v in 1 0; do echo "== bench x3 LN_FUSION=$v"; DVD_LN_FUSION=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_lnf$v.txt 2>&1; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_lnf$v.txt').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline']['achieved'],d['kernel_share'])"; done
for v in 0 1; do echo "== bench x3 docs=16 LN_FUSION=$v"; DVD_LN_FUSION=$v timeout 300 python bench.py --steps 5 --warmup 3 --docs 16 --no-extras --no-cpu-baseline > gpurun_out/bench16_lnf$v.txt 2>&1; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench16_lnf$v.txt').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline']['achieved'],d['clocks'])"; done
echo "== bench bf16"; timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16 --no-extras --no-cpu-baseline > gpurun_out/bench_bf16.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_bf16.txt').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline']['achieved'])"

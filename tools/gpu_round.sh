#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_all.txt
echo "== gemm graph a16w3"; timeout 300 python tools/gemm_bench.py --graph --a16w3 2>&1 | tail -10
for v in 0 1; do echo "== bench x3 QKV_3PASS=$v"; DVD_QKV_3PASS=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_q$v.txt 2>&1; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_q$v.txt').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline'],d['kernel_share'],d.get('parity'))"; done
echo "== bench full (parity)"; timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench_par.txt 2>&1; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_par.txt').read().strip().splitlines()[-1]);print(d['value'],d.get('parity'),d.get('cpu_baseline'))"

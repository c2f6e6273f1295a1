#!/bin/bash
# One gpurun call: kernel checks first (cheap, bounded), then the GPU test suite, then a short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gemm_bench v1" ; DVD_GEMM_V1=1 timeout 300 python tools/gemm_bench.py --check > gpurun_out/gemm_v1.txt 2>&1; echo "rc=$?"; tail -22 gpurun_out/gemm_v1.txt
echo "== gemm_bench pair" ; timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_pair.txt 2>&1; echo "rc=$?"; tail -22 gpurun_out/gemm_pair.txt
echo "== pytest" ; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.txt 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest.txt
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench.txt

#!/bin/bash
mkdir -p gpurun_out
for bn in 256 192 128; do for sp in 1 2 3; do echo "== bn=$bn sp=$sp"; DVD_GEMM_BN=$bn DVD_GEMM_SPLITS=$sp timeout 200 python tools/gemm_bench.py --graph 2>&1 | grep -v "x8 docs" ; done; done > gpurun_out/gemm_sweep.txt 2>&1
cat gpurun_out/gemm_sweep.txt

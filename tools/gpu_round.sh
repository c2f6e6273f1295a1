#!/bin/bash
# Evidence round: extra bench configs, ncu --set full of the dominant kernels, compute-sanitizer.
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu-baseline"
: > gpurun_out/r2_bench_lines.jsonl
run() { echo "== $*"; timeout 900 "$@" 2>gpurun_out/last_err.txt | tail -1 >> gpurun_out/r2_bench_lines.jsonl; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_lines.jsonl | cut -c1-200; }
run $B --steps 10 --warmup 3 --docs 16
run $B --steps 5 --warmup 3 --docs 64
run $B --steps 10 --warmup 3 --height 4032 --width 3024
run $B --steps 10 --warmup 3 --height 4032 --width 3024 --docs 16
run $B --steps 10 --warmup 3 --docs 16 --precision bf16
for S in 10 100 1000; do for D in 1 8; do st=5; [ $S -ge 100 ] && st=2; run $B --steps $st --warmup 3 --diffusion-steps $S --docs $D; done; done
echo "== ncu tensor kernels (in-pipeline, graphs off)"
DVD_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_pair|k_attn_tc" -s 60 -c 14 -o gpurun_out/r2_tensor python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_tensor.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r2_tensor.ncu-rep > gpurun_out/r2_ncu_tensor_kernels.txt 2>&1; grep -E "^void|tensor_cycles|time_duration" gpurun_out/r2_ncu_tensor_kernels.txt | head -60
echo "== ncu unwarp"
timeout 600 ncu --set full --clock-control none -k regex:"k_unwarp" -s 4 -c 4 -o gpurun_out/r2_unwarp env UW_LAUNCHES=2 python tools/unwarp_bench.py > gpurun_out/ncu_unwarp.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r2_unwarp.ncu-rep > gpurun_out/r2_ncu_unwarp.txt 2>&1; grep -E "^void|dram__bytes|time_duration" gpurun_out/r2_ncu_unwarp.txt | head -40
echo "== unwarp bench"; timeout 300 python tools/unwarp_bench.py > gpurun_out/r2_unwarp_bench.txt 2>&1; cat gpurun_out/r2_unwarp_bench.txt
echo "== sanitizer memcheck"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_gemm_bf16x3_tcgen05 or test_gemm_bf16_tcgen05 or test_attention_fp16 or test_attention_bf16 or test_unwarp_matches_reference_golden or test_unwarp_tma_batched" > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_sanitizer_memcheck.txt
echo "== sanitizer racecheck"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_gemm_bf16x3_tcgen05 and (128-128-64 or 512-64-64 or 384-192-128)" > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck.txt
echo "== sanitizer synccheck"
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_gemm_bf16x3_tcgen05 and (128-128-64 or 512-64-64) or test_attention_fp16" > gpurun_out/r2_sanitizer_synccheck.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_sanitizer_synccheck.txt

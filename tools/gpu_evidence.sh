#!/bin/bash
# Produces the round's evidence files on a B200 box (run from the repo root, e.g. `gpurun --timeout 3300 -- 'bash tools/gpu_evidence.sh'`);
# everything lands in gpurun_out/ (keep it under 64 MiB: the ncu report stays in /tmp), the keepers are copied to profiles/ by hand.
#   one GPU : tests, the default bench line with every leg, the reference arm, extra configs, launch list, ncu --set full of the tensor
#             kernels inside a step, micro-benchmarks, traces (instrumented build), compute-sanitizer
#   8 GPUs  : `gpurun --gpus 8 -- 'bash tools/gpu_multi.sh'`
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.txt
echo "== bench default (all legs)"; timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default_line.json 2>gpurun_out/bench_default_err.txt; echo "rc=$?"; cut -c1-300 gpurun_out/r2_bench_default_line.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_line.json 2>gpurun_out/bench_ref_err.txt; echo "rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference_line.json
B="python bench.py --no-extras --no-cpu-baseline"
: > gpurun_out/r2_bench_lines.jsonl
run() { echo "== $*"; timeout 900 "$@" 2>gpurun_out/last_err.txt | tail -1 >> gpurun_out/r2_bench_lines.jsonl; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_lines.jsonl | cut -c1-120; }
run $B --steps 10 --warmup 3 --docs 16
run $B --steps 5 --warmup 3 --docs 64
run $B --steps 10 --warmup 3 --height 4032 --width 3024
run $B --steps 10 --warmup 3 --height 4032 --width 3024 --docs 16
run $B --steps 10 --warmup 3 --docs 16 --precision bf16
run $B --steps 5 --warmup 3 --diffusion-steps 10 --docs 1
run $B --steps 5 --warmup 3 --diffusion-steps 10 --docs 8
run $B --steps 2 --warmup 3 --diffusion-steps 100 --docs 1
run $B --steps 2 --warmup 3 --diffusion-steps 100 --docs 8
run $B --steps 2 --warmup 3 --diffusion-steps 1000 --docs 1
echo "== graph trace"; timeout 300 python tools/graph_trace.py --out gpurun_out/r2_graph_trace.txt > /dev/null 2>&1; echo "rc=$?"; head -12 gpurun_out/r2_graph_trace.txt
echo "== micro-benchmarks"
timeout 300 python tools/gemm_bench.py --graph > gpurun_out/r2_gemm_graph_pair.txt 2>&1; timeout 300 python tools/gemm_bench.py --graph --a16w3 >> gpurun_out/r2_gemm_graph_pair.txt 2>&1
timeout 300 python tools/attn_bench.py 2>&1 | grep k_attn > gpurun_out/r2_attn_bench.txt; DVD_ATTN_V1=1 timeout 300 python tools/attn_bench.py --d 256 2>&1 | grep k_attn >> gpurun_out/r2_attn_bench.txt; cat gpurun_out/r2_attn_bench.txt
if [ -f dvd_b200/libdvd_b200_trace.so ]; then      # DVD_NVCC_EXTRA="-DDVD_GEMM_TRACE -DDVD_GEMM_TRACE2 -DDVD_ATTN_TRACE" python -m dvd_b200.build
  echo "== traces"
  DVD_LIB=dvd_b200/libdvd_b200_trace.so timeout 200 python tools/attn_bench.py --d 256 --trace 2>&1 | tail -10 > gpurun_out/r2_attn_trace.txt
  DVD_LIB=dvd_b200/libdvd_b200_trace.so DVD_NO_GRAPH=1 timeout 300 python tools/step_trace.py 2048 1536 4608 > gpurun_out/r2_step_trace.txt 2>&1; tail -14 gpurun_out/r2_step_trace.txt
fi
echo "== ncu launch list"
DVD_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bf16x3_batch1.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_bf16x3_batch1.csv > gpurun_out/r2_launches_bf16x3_batch1.txt 2>&1; head -5 gpurun_out/r2_launches_bf16x3_batch1.txt
echo "== ncu tensor kernels (in-pipeline, graphs off)"
DVD_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none -k regex:"k_gemm_pair|k_attn_tc|k_attn_pair" -s 60 -c 10 -o /tmp/r2_tensor python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_tensor.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py /tmp/r2_tensor.ncu-rep > gpurun_out/r2_ncu_tensor_kernels.txt 2>&1; grep -c "^void" gpurun_out/r2_ncu_tensor_kernels.txt
echo "== sanitizer memcheck"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_gemm_bf16x3_tcgen05 or test_gemm_bf16_tcgen05 or test_gemm_fp16_activation or test_attention_fp16 or test_attention_bf16 or test_attention_d256_single or test_unwarp_matches_reference_golden or test_unwarp_tma_batched" > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.txt
echo "== sanitizer racecheck"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_gemm_bf16x3_tcgen05 and (128-128-64 or 512-64-64 or 384-192-128) or test_gemm_fp16_activation and 256-64-64" > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.txt
echo "== sanitizer synccheck"
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_gemm_bf16x3_tcgen05 and (128-128-64 or 512-64-64) or test_gemm_fp16_activation and 256-64-64 or test_attention_fp16 or test_attention_d256_single" > gpurun_out/r2_sanitizer_synccheck.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_sanitizer_synccheck.txt
du -sh gpurun_out

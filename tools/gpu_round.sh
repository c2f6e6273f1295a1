#!/bin/bash
mkdir -p gpurun_out
echo "== pytest attention + parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "attention or stages or sampling_bf16x3 or psnr" > gpurun_out/pytest_attn.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_attn.txt
echo "== sanitizer synccheck"
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_gemm_bf16x3_tcgen05 and (128-128-64 or 512-64-64) or test_gemm_fp16_activation and 256-64-64 or test_attention_fp16 or test_attention_d256_single" > gpurun_out/r2_sanitizer_synccheck.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_sanitizer_synccheck.txt
echo "== sanitizer memcheck attention"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_attention_fp16 or test_attention_bf16 or test_attention_d256_single" > gpurun_out/r2_sanitizer_memcheck_attn.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck_attn.txt
echo "== attn bench"; timeout 200 python tools/attn_bench.py --d 256 2>&1 | grep k_attn
echo "== bench x3"; timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_x3.txt 2>&1; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_x3.txt').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline']['achieved'],d['roofline']['traffic'])"

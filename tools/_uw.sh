cd /root/repo
DVD_GEMM_V1=1 timeout 300 python tools/gemm_trace.py > gpurun_out/r28_gemm_trace.txt 2>&1

cd /root/repo
for bn in 128 256; do echo V3 BN=$bn; DVD_GEMM_V3=1 DVD_GEMM_BN=$bn timeout 300 python tools/gemm_bench.py 2>&1 | tail -10; done

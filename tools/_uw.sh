cd /root/repo
DVD_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:"k_gemm_tc|k_attn_tc" --launch-skip 150 -c 12 -f -o gpurun_out/r31_tensor python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r31_ncu.log 2>&1
ls -la gpurun_out/

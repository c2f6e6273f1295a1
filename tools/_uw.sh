cd /root/repo
for mb in 2 3; do echo NS=3 MINB=$mb; DVD_UNWARP_MINB=$mb UW_AMP=0.005 timeout 300 python tools/unwarp_bench.py; done

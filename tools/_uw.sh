cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k unwarp 2>&1 | tail -5
for mb in 3 4; do for amp in 0.02 0.005; do echo MINB=$mb; DVD_UNWARP_MINB=$mb UW_AMP=$amp timeout 300 python tools/unwarp_bench.py; done; done

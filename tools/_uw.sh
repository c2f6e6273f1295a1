cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k unwarp 2>&1 | tail -3
for amp in 0.05 0.02 0.005; do UW_AMP=$amp timeout 300 python tools/unwarp_bench.py 2>&1 | grep f32; done
python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], json.dumps(d['roofline_unwarp']['random_init_map']))"

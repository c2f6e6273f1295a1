cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu -k "gemm_bf16 or static or sampling_bf16 or batch" 2>&1 | tail -2
python bench.py --docs 64 --steps 3 --no-cpu-baseline 2>gpurun_out/r32_docs64.err | tail -1 > gpurun_out/r32_docs64.json; tail -3 gpurun_out/r32_docs64.err | cut -c1-300; python -c "
import json
d=json.loads(open('gpurun_out/r32_docs64.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['attention_tflops'])"

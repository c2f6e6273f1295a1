cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/r30_b1.json 2>gpurun_out/r30_b1.err; tail -1 gpurun_out/r30_b1.json | cut -c1-160
python bench.py --precision fp32 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r30_fp32.json; cut -c1-160 gpurun_out/r30_fp32.json
python bench.py --docs 16 --steps 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r30_docs16.json; cut -c1-160 gpurun_out/r30_docs16.json
python bench.py --height 4032 --width 3024 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r30_4032.json; cut -c1-160 gpurun_out/r30_4032.json
python bench.py --height 4032 --width 3024 --docs 16 --steps 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r30_4032_docs16.json; cut -c1-160 gpurun_out/r30_4032_docs16.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r30_ref.json; cut -c1-200 gpurun_out/r30_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

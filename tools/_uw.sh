cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu -k "sampling or stage or batch or dropin" 2>&1 | tail -3
python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-250
timeout 600 python tools/graph_trace.py --out gpurun_out/r23_graph_trace.txt 2>&1 | grep "dwconv\|one document"

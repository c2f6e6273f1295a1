cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu -k "tma_batched" 2>&1 | tail -15

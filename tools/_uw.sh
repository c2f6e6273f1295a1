cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu -k "two_chain" 2>&1 | tail -15

timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tests/_full_probe.py 64 3 64 2>&1 | tail -5
SA_GPU_COARSE_BLOCKED_MIN=100000 timeout 300 python tests/_full_probe.py 64 3 64 2>&1 | tail -2

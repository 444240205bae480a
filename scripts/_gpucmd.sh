timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tests/_full_probe.py 128 4 64 2>&1 | grep -E "ml_build|^\{|pcg iters" | cut -c1-900

timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tests/_bench_probe.py 128 52 2 2>&1 | grep "resident step" | tail -1

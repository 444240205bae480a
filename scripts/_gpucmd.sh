SA_GPU_PIPE_DEBUG=1 timeout 300 python tests/_bench_probe.py 128 52 2 2>&1 | grep "e2e\|pipe\|level_create\|noprof" | tail -9
timeout 300 python tests/_bench_probe.py 128 52 2 2>&1 | grep "e2e" | tail -4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2

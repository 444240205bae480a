SA_GPU_MAX_CLASS=2 timeout 300 python tests/_bench_probe.py 128 52 2 2>&1 | grep "resident step" | tail -1
SA_GPU_MAX_CLASS=1 timeout 300 python tests/_bench_probe.py 128 52 2 2>&1 | grep "resident step" | tail -1

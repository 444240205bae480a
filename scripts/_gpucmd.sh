timeout 600 python scripts/_q2probe.py 32 3 2>&1 | tail -6

timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01d.log 2> gpurun_out/bench_r01d.err; tail -c 300 gpurun_out/bench_r01d.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r01d.log') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print(d['e2e']); print(d['hierarchy']['setup_s'], d['hierarchy']['pcg_s'])
print(d['cpu_baseline'], d['cpu_baseline_1core'])
PY

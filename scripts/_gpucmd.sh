timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01c.log 2> gpurun_out/bench_r01c.err; tail -c 300 gpurun_out/bench_r01c.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r01c.log') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print(d['e2e']); print(d['roofline']['frac'], d['roofline']['stage_ms'], d['roofline']['traffic'])
print(d['hierarchy']['setup_s'], d['hierarchy']['pcg_s'], d['hierarchy']['pcg_iterations'], d['hierarchy']['stage_s'])
print(d['roofline_spmv']['frac'], d['roofline_smoother']['frac'], d['cpu_baseline'])
PY

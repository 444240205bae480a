TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/mgpu_hierarchy_check.py 64 3 2>&1 | grep -E "PASS|FAIL|iters|setup" | head -6
timeout 600 $TR tests/mgpu_solve_check.py 128 4 2>&1 | grep -E "PASS|FAIL|iters" | head -4
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_b.log 2> gpurun_out/bench_2gpu_b.err; tail -c 200 gpurun_out/bench_2gpu_b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu_b.log') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['hierarchy']['setup_s'], d['hierarchy']['pcg_iterations'], {k:v for k,v in d['hierarchy']['stage_s'].items() if v>0.05})
PY

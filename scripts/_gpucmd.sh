TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.log 2> gpurun_out/bench_4gpu.err; tail -c 300 gpurun_out/bench_4gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_4gpu.log') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['hierarchy']['setup_s'], d['hierarchy']['pcg_iterations'], {k:v for k,v in d['hierarchy']['stage_s'].items() if v>0.05})
PY
timeout 300 $TR bench.py --impl reference --gpus 4 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300

SA_GPU_ALLOC_DEBUG=1 timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu 2> gpurun_out/alloc.err > gpurun_out/alloc.log
grep alloc gpurun_out/alloc.err | awk '{s+=$5; n++} END {print n, "allocs >=64MB, total ms", s}'
grep alloc gpurun_out/alloc.err | sort -k5 -n -r | head -4
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/alloc.log') if l.startswith('{')][0])
print(d['value'], d['e2e']['value'], d['hierarchy']['setup_s'], {k:v for k,v in d['hierarchy']['stage_s'].items() if v>0.02})
PY

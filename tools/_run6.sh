mkdir -p gpurun_out
MOGP_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/s17_scale_8_trace.json 2> gpurun_out/s17_scale_8_trace.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29509 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s17_scale_8.json 2> gpurun_out/s17_scale_8.err
python - <<'PY'
import json
for f in ['gpurun_out/s17_scale_8.json','gpurun_out/s17_scale_8_trace.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['e2e']['value'], json.dumps(d['phases_ms_per_step']), json.dumps(d['host_wall_ms_per_step']), json.dumps(d.get('cholesky')), d.get('parity_vs_cpu_sample'))
PY
grep "mogp trace" gpurun_out/s17_scale_8_trace.err | tail -60

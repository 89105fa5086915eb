#!/bin/bash
# N-GPU runs of the C3 bench on one box (run under gpurun --gpus 8): one JSON line per N in gpurun_out/r02_scale_<N>.json,
# then C5 (256 outputs x n=8192) on all GPUs.  usage: tools/scale_run.sh "4 8" [c5]
mkdir -p gpurun_out
for n in ${1:-1 2 4 8}; do
  if [ $n = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-other > gpurun_out/r02_scale_1.json 2> gpurun_out/r02_scale_1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_scale_$n.json 2> gpurun_out/r02_scale_$n.err
  fi
done
if [ "$2" = "c5" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus 8 --workload c5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_c5_8.json 2> gpurun_out/r02_c5_8.err
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_scale_*.json")) + glob.glob("gpurun_out/r02_c5_8.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "value %.4f" % d["value"], "e2e %.4f" % d["e2e"]["value"], d["phases_ms_per_step"], d["host_wall_ms_per_step"],
              d["roofline"].get("frac"), d["roofline"].get("fp64_fallbacks"), d.get("parity_vs_cpu_sample"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/r02_scale_8.err gpurun_out/r02_c5_8.err 2>/dev/null

timeout 300 python tools/solve_sweep.py 2>&1 | tail -6
timeout 300 python bench.py --workload c4 --no-other --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', d['value'], d['phases_ms_per_step'])"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2

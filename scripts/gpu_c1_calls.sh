# usage: bash scripts/gpu_c1_calls.sh  - c1 bench (no cpu leg): value + per-call event times of the extraction calls
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); pc=d['roofline']['per_call_ms']
print('c1', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {k: round(v*1e3,1) for k,v in pc.items() if 'mt_' in k or 'bones' in k})"

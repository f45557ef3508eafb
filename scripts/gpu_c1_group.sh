# usage: bash scripts/gpu_c1_group.sh  - c1 bench (no cpu leg): value + the raster-backward group's event times
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('c1', round(d['value'],1), round(d['ms_per_step'],4), 'group us', round(r['us_per_launch'],2), 'frac', round(r['frac'],4), {k: round(v['ms']*1e3,2) for k,v in r['kernels'].items()})"

# usage: bash scripts/gpu_configs.sh TAG  - bench lines of the other BASELINE configs (c2 bird, c3 fauna, c4 visualize)
T=$1
mkdir -p gpurun_out
for C in c2 c3 c4; do
  python bench.py --config $C --steps 50 --warmup 5 > gpurun_out/bench_${C}_$T.json 2> gpurun_out/bench_${C}_$T.err; tail -2 gpurun_out/bench_${C}_$T.err | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/bench_${C}_$T.json')); print('$C', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'roof', d['roofline']['frac'], d['roofline']['us_per_launch'], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), d['gpu_launches']); print({k: (round(v['ms']*1e3,1), round(v['frac'],3)) for k,v in d['roofline']['kernels'].items()})"
done

mkdir -p gpurun_out
for m in fp32 tf32 fp16; do
python bench.py --steps 20 --warmup 3 --mlps --mlp-math $m --no-cpu > gpurun_out/bench_mlps_$m.json 2> gpurun_out/bench_mlps_$m.err; tail -1 gpurun_out/bench_mlps_$m.err | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/bench_mlps_$m.json')); print('M1b $m', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
done
python -m pytest tests -m gpu -q -x -k "pair or antialias" 2>&1 | tail -1

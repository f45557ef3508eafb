mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "skinning or lbs or hot_path" > gpurun_out/test_ab2.log 2>&1; tail -2 gpurun_out/test_ab2.log | cut -c1-300
for i in 1 2; do
for v in on off; do
python bench.py --steps 100 --warmup 5 --no-cpu --autograd-threads $v > gpurun_out/ab2_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/ab2_$v.json')); print('threads=$v', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'])"
done; done

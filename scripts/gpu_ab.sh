# A/B on one box: bench with and without the fused AA pair, twice each, interleaved
mkdir -p gpurun_out
for i in 1 2; do
for v in 1 0; do
B2A_FUSE_AA_PAIR=$v python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/ab_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('pair=$v', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['clocks'])"
done; done
nproc; lscpu | grep -i "model name\|MHz" | head -4

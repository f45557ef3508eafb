# usage: bash scripts/gpu_n2.sh TAG [N]   - N=1 and N=N (default 2) bench on one box, scaling efficiency
T=$1; N=${2:-2}
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_n1_$T.json 2>/dev/null
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_n${N}_$T.json 2> gpurun_out/bench_n${N}_$T.err; tail -2 gpurun_out/bench_n${N}_$T.err | cut -c1-300
python -c "
import json
a=json.load(open('gpurun_out/bench_n1_$T.json')); b=json.load(open('gpurun_out/bench_n${N}_$T.json'))
print('N1', round(a['value']), a['ms_per_step'], round(a['e2e']['value'])); print('N$N', round(b['value']), b['ms_per_step'], round(b['e2e']['value']), b['run']['allreduce_bytes_per_step'], 'eff', b['value']/$N/a['value'], b['e2e']['value']/$N/a['e2e']['value'])"

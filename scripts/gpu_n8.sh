mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/n$N.json 2>gpurun_out/n$N.err; tail -2 gpurun_out/n$N.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/n$N.json')); print('N=$N', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['clocks'])"
nproc

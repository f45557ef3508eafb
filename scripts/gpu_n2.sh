mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/n1.json 2>gpurun_out/n1.err; python -c "
import json; d=json.load(open('gpurun_out/n1.json')); print('N=1', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/n2.json 2>gpurun_out/n2.err; tail -2 gpurun_out/n2.err; python -c "
import json; d=json.load(open('gpurun_out/n2.json')); print('N=2', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'])"
nproc

mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r1p.log 2>&1; tail -2 gpurun_out/smoke_r1p.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench2_r1p.json 2> gpurun_out/bench2_r1p.err; tail -3 gpurun_out/bench2_r1p.err; python -c "
import json; d=json.load(open('gpurun_out/bench2_r1p.json')); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])"
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err; tail -3 gpurun_out/bench_r1p.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1p.json')); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
python bench.py --steps 10 --warmup 3 --mlps --no-cpu > gpurun_out/bench_mlps_r1p.json 2> gpurun_out/bench_mlps_r1p.err; tail -3 gpurun_out/bench_mlps_r1p.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mlps_r1p.json')); print('M1b', d['value'], d['ms_per_step'], d['e2e']['value'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1p.json 2> gpurun_out/bench_ref_r1p.err; tail -3 gpurun_out/bench_ref_r1p.err; cat gpurun_out/bench_ref_r1p.json | cut -c1-300

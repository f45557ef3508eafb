# usage: bash scripts/gpu_e2e_segments.sh N  - host segments of the e2e harness (rank 0) at N=1 and N=N on the same box
N=${1:-4}
B2A_E2E_PROFILE=1 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu 2> gpurun_out/e2e_seg_n1.err > gpurun_out/e2e_seg_n1.json; grep "e2e host segments" gpurun_out/e2e_seg_n1.err | cut -c1-400
B2A_E2E_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu 2> gpurun_out/e2e_seg_n$N.err > gpurun_out/e2e_seg_n$N.json; grep "e2e host segments" gpurun_out/e2e_seg_n$N.err | cut -c1-400
python -c "
import json
a=json.load(open('gpurun_out/e2e_seg_n1.json')); b=json.load(open('gpurun_out/e2e_seg_n$N.json'))
print('N1 value %.0f e2e %.0f (%.3f ms) | N$N value %.0f e2e %.0f (%.3f ms)' % (a['value'], a['e2e']['value'], a['e2e']['ms_per_step'], b['value'], b['e2e']['value'], b['e2e']['ms_per_step']))"

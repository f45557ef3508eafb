mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/test_r1k.log 2>&1; tail -15 gpurun_out/test_r1k.log
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; tail -3 gpurun_out/bench_r1k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1k.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['roofline']['raster_backward_group'])); print(json.dumps(d['roofline']['per_call_ms']))"
B2A_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'aa_|gb_|eb_|raster_|lbs_|mt_|normals_|xfm_|adj_' -c 70 -o gpurun_out/prof_r1k -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_r1k.log 2>&1; tail -2 gpurun_out/ncu_full_r1k.log

mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_r1t.log 2>&1; tail -5 gpurun_out/test_r1t.log | cut -c1-400
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_r1t.json 2> gpurun_out/bench_r1t.err; tail -3 gpurun_out/bench_r1t.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1t.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac']); print(json.dumps(d['roofline']['raster_backward_group'])); print(json.dumps(d['roofline']['per_call_ms']))"
python bench.py --steps 10 --warmup 3 --mlps --no-cpu > gpurun_out/bench_mlps_r1t.json 2> gpurun_out/bench_mlps_r1t.err; tail -3 gpurun_out/bench_mlps_r1t.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mlps_r1t.json')); print('M1b', d['value'], d['ms_per_step'], d['e2e']['value'])"
B2A_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'aa_|gb_|raster_|lbs_|mt_|normals_|xfm_|adj_|eb_' -c 70 -o gpurun_out/prof_r1t -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_r1t.log 2>&1; tail -2 gpurun_out/ncu_full_r1t.log

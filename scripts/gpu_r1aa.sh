mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_r1aa.log 2>&1; tail -3 gpurun_out/test_r1aa.log | cut -c1-800
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r1aa.json 2> gpurun_out/bench_r1aa.err; tail -2 gpurun_out/bench_r1aa.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1aa.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['gpu_launches'], d['cpu_baseline']); print(json.dumps(d['roofline']['raster_backward_group'])[:700]); print(json.dumps(d['roofline']['per_call_ms']))"
python bench.py --steps 20 --warmup 3 --mlps --no-cpu > gpurun_out/bench_mlps_r1aa.json 2> gpurun_out/bench_mlps_r1aa.err; tail -2 gpurun_out/bench_mlps_r1aa.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mlps_r1aa.json')); print('M1b', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
B2A_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1aa.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_r1aa.log 2>&1; tail -1 gpurun_out/ncu_list_r1aa.log | cut -c1-300
B2A_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'aa_|gb_|raster_|lbs_|mt_|normals_|xfm_|adj_|eb_|shade_|af_' -c 60 -o gpurun_out/prof_r1aa -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_r1aa.log 2>&1; tail -2 gpurun_out/ncu_full_r1aa.log | cut -c1-300
ls -la gpurun_out/prof_r1aa.ncu-rep

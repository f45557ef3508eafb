# usage: bash scripts/gpu_full.sh TAG   - full validation + profile refresh (tests, bench, M1b, launch list, ncu --set full)
T=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_$T.log 2>&1; tail -3 gpurun_out/test_$T.log | cut -c1-400
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -2 gpurun_out/bench_$T.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_$T.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['traffic'], d['gpu_launches']); print(json.dumps(d['cpu_baseline'])[:900]); print(json.dumps(d['roofline']['raster_backward_group'])[:700]); print(json.dumps(d['roofline']['per_call_ms']))"
python bench.py --steps 20 --warmup 3 --mlps --no-cpu > gpurun_out/bench_mlps_$T.json 2> gpurun_out/bench_mlps_$T.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mlps_$T.json')); print('M1b', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$T.json
python profiles/host_segments.py 100 > gpurun_out/host_segments_$T.txt 2>&1; head -4 gpurun_out/host_segments_$T.txt
python profiles/timeline.py 5 > gpurun_out/timeline_$T.txt 2>&1; sed -n 3,5p gpurun_out/timeline_$T.txt | cut -c1-200
B2A_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_$T.log 2>&1; tail -1 gpurun_out/ncu_list_$T.log | cut -c1-200
B2A_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'aa_|gb_|raster_|lbs_|mt_|normals_|xfm_|adj_|eb_|shade_|af_' -c 60 -o gpurun_out/prof_$T -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$T.log 2>&1; tail -1 gpurun_out/ncu_full_$T.log | cut -c1-200
ls -la gpurun_out/prof_$T.ncu-rep

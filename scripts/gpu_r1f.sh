mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/test_r1f.log 2>&1; tail -3 gpurun_out/test_r1f.log
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['roofline']['raster_backward_group'])); print(json.dumps(d['roofline']['per_call_ms']))"
B2A_PROFILE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'aa_' -c 12 -o gpurun_out/prof_r1f -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_r1f.log 2>&1; tail -2 gpurun_out/ncu_full_r1f.log
python profiles/host_breakdown.py 30 > gpurun_out/host_r1f.txt 2>&1; tail -3 gpurun_out/host_r1f.txt

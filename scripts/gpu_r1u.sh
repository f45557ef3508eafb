mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_r1u.log 2>&1; tail -5 gpurun_out/test_r1u.log | cut -c1-400
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err; tail -3 gpurun_out/bench_r1u.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1u.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline']); print(json.dumps(d['roofline']['raster_backward_group'])); print(json.dumps(d['roofline']['per_call_ms']))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1u.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_r1u.log 2>&1; tail -2 gpurun_out/ncu_list_r1u.log
python profiles/host_breakdown.py 30 > gpurun_out/host_r1u.txt 2>&1; tail -3 gpurun_out/host_r1u.txt

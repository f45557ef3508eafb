mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_r1q.log 2>&1; tail -5 gpurun_out/test_r1q.log | cut -c1-400
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; tail -3 gpurun_out/bench_r1q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1q.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac']); print(json.dumps(d['roofline']['raster_backward_group'])); print(json.dumps(d['roofline']['per_call_ms']))"
python bench.py --steps 10 --warmup 3 --mlps --no-cpu > gpurun_out/bench_mlps_r1q.json 2> gpurun_out/bench_mlps_r1q.err; tail -3 gpurun_out/bench_mlps_r1q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mlps_r1q.json')); print('M1b', d['value'], d['ms_per_step'], d['e2e']['value'])"

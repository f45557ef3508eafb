mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_r1z.log 2>&1; tail -5 gpurun_out/test_r1z.log | cut -c1-800
python profiles/host_segments.py 50 > gpurun_out/host_segments_r1z.txt 2>&1; cat gpurun_out/host_segments_r1z.txt | cut -c1-150
for i in 1 2; do python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; tail -2 gpurun_out/bench_r1z.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1z.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['gpu_launches']); print(json.dumps(d['roofline']['raster_backward_group'])[:600])"; done
python profiles/timeline.py 5 > gpurun_out/timeline_r1z.txt 2>&1; head -45 gpurun_out/timeline_r1z.txt | cut -c1-150

mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "pair or shade or light or analytic or marching_tets_golden or lbs or skinning_golden or render_mesh_matches" > gpurun_out/sanitizer_r1ab.log 2>&1; echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitizer_r1ab.log | head -8
for i in 1 2; do python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_r1ab.json 2> gpurun_out/bench_r1ab.err; tail -1 gpurun_out/bench_r1ab.err | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/bench_r1ab.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['gpu_launches'])"; done
python profiles/host_segments.py 50 > gpurun_out/host_segments_r1ab.txt 2>&1; head -12 gpurun_out/host_segments_r1ab.txt

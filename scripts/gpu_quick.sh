# usage: bash scripts/gpu_quick.sh TAG [pytest -k expr]
T=$1; K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/test_$T.log 2>&1; else python -m pytest tests -m gpu -q > gpurun_out/test_$T.log 2>&1; fi
tail -5 gpurun_out/test_$T.log | cut -c1-600
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$T.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['gpu_launches']); print(json.dumps(d['roofline']['raster_backward_group'])); print(json.dumps(d['roofline']['per_call_ms']))"

for v in 0 1 2 0 1 2; do
B2A_AA_VAR=$v python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/aavar.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/aavar.json')); k=d['roofline']['raster_backward_group']['kernels']['aa_bwd_pair']; print('B2A_AA_VAR=$v', round(k['ms']*1e3,2), 'us', round(k['frac'],3), round(d['value']))"
done
python -m pytest tests -m gpu -q -x -k "pair" 2>&1 | tail -1
B2A_AA_VAR=2 python -m pytest tests -m gpu -q -x -k "pair or render_mesh_matches or hot_path" 2>&1 | tail -1

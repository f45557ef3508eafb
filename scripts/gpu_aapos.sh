mkdir -p gpurun_out
for v in "592" "592,last" "148" "148,last" "37,last" "1184"; do
B2A_AA_POS="$v" python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/aapos.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/aapos.json')); k=d['roofline']['raster_backward_group']['kernels']['aa_bwd_pair']; print('B2A_AA_POS=$v', round(k['ms']*1e3,2), 'us', round(k['frac'],3), round(d['value']))"
done

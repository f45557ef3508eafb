# usage: bash scripts/gpu_final.sh TAG  - the round's final single-GPU measurement pass (every command under its own timeout)
T=$1
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > $O/pytest_gpu_$T.txt; cat $O/pytest_gpu_$T.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 5 > $O/bench_c1_$T.json 2> $O/bench_c1_$T.err; tail -1 $O/bench_c1_$T.err | cut -c1-200
for C in c2 c3 c4; do
  timeout 400 python bench.py --config $C --steps 50 --warmup 5 > $O/bench_${C}_$T.json 2> $O/bench_${C}_$T.err; tail -1 $O/bench_${C}_$T.err | cut -c1-200
done
timeout 400 python bench.py --mlps --steps 30 --warmup 3 --no-cpu > $O/bench_mlps_$T.json 2> $O/bench_mlps_$T.err; tail -1 $O/bench_mlps_$T.err | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref_$T.json 2> $O/bench_ref_$T.err; tail -1 $O/bench_ref_$T.err | cut -c1-200
timeout 300 python profiles/host_profile.py > $O/host_profile_$T.txt 2>&1; tail -40 $O/host_profile_$T.txt | head -5
timeout 300 python profiles/timeline.py > $O/timeline_$T.txt 2>&1; tail -3 $O/timeline_$T.txt
B2A_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_list_$T.log 2>&1
python profiles/ncu_summary.py $O/launches_$T.csv 2 > $O/launches_$T.txt 2>&1
python - <<PY
import json
for c in ("c1","c2","c3","c4","mlps","ref"):
    try:
        d=json.loads(open("$O/bench_%s_$T.json"%c).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(c, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roof", r.get("frac"), r.get("us_per_launch"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(c, "FAILED", e)
PY

# usage: bash scripts/gpu_list.sh TAG   - per-kernel device times of one steady-state step (ncu launch list; cold-cache, serialised)
T=$1
mkdir -p gpurun_out
B2A_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_$T.log 2>&1
python profiles/ncu_summary.py gpurun_out/launches_$T.csv 2 > gpurun_out/launches_$T.txt 2>&1; head -70 gpurun_out/launches_$T.txt

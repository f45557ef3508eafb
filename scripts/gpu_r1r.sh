mkdir -p gpurun_out
python profiles/host_breakdown.py 40 > gpurun_out/host_r1r.txt 2>&1; tail -3 gpurun_out/host_r1r.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1r.csv env B2A_PROFILE=1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_r1r.log 2>&1; tail -1 gpurun_out/ncu_launch_r1r.log | cut -c1-200

mkdir -p gpurun_out
python profiles/timeline.py 5 > gpurun_out/timeline_r1v.txt 2>&1; head -60 gpurun_out/timeline_r1v.txt | cut -c1-180

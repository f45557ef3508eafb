mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "captured" > gpurun_out/test_r1af.log 2>&1; tail -15 gpurun_out/test_r1af.log | cut -c1-300
timeout 500 python profiles/c4_step.py > gpurun_out/c4_step_r1.txt 2>&1; grep -A1 "^==" gpurun_out/c4_step_r1.txt | cut -c1-200; tail -5 gpurun_out/c4_step_r1.txt | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -2

mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_r1ag.log 2>&1; tail -12 gpurun_out/test_r1ag.log | cut -c1-400
timeout 500 python profiles/c4_step.py > gpurun_out/c4_step_r2.txt 2>&1; grep -A1 "^==" gpurun_out/c4_step_r2.txt | cut -c1-200; sed -n "/device activities per frame/,\$p" gpurun_out/c4_step_r2.txt | head -16 | cut -c1-150

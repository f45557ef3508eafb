timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-2} --master-addr 127.0.0.1 --master-port 29517 scripts/p2p_check.py > gpurun_out/p2p_check.txt 2>&1; echo rc=$?; grep -v "^W1\|^\*\*\*\|OMP" gpurun_out/p2p_check.txt | tail -15
bash scripts/gpu_n2.sh r2s_n${1:-2} ${1:-2}

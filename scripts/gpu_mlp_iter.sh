timeout 120 python scripts/dbg_mlp_speed.py 2>&1 | tail -14
timeout 400 python -m pytest tests/test_gpu_field_mlp.py -x -q -m gpu 2>&1 | tail -4

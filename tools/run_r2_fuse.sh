python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused" 2>&1 | tail -6
python tools/time_step.py 20000 4 unfused 2>&1 | tail -1
python tools/time_step.py 20000 4 2>&1 | tail -1
python tools/time_step.py 10000 4 unfused 2>&1 | tail -1
python tools/time_step.py 10000 4 2>&1 | tail -1
python tools/microbench_r2.py 20000 2>&1 | grep -i "gram" | head -4

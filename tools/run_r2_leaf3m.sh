python tools/bench_leaf.py 2>&1 | tail -14
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
python tools/time_chol.py 20000,10000,4096,2048,1024 2>&1 | tail -1
python tools/time_step.py 20000 3 2>&1 | tail -1
LGP_EARLY_INVERSE=1 python tools/time_step.py 20000 3 2>&1 | tail -1
python tools/time_step.py 10000 4 2>&1 | tail -1
LGP_EARLY_INVERSE=1 python tools/time_step.py 10000 4 2>&1 | tail -1
python tools/c1_breakdown.py 2>&1 | head -4

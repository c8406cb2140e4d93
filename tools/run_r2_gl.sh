python tools/bench_gemm.py 2>&1 | tail -6
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "dgemm or chol" 2>&1 | tail -2
python tools/time_chol.py 20000,10000 2>&1 | tail -1
python tools/time_step.py 20000 3 2>&1 | tail -1
python tools/time_step.py 10000 4 2>&1 | tail -1

python tools/time_solve.py 1000,4096,20000 2>&1 | tail -3
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_reference_vectors.py -m gpu -q -x 2>&1 | tail -3
python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k "single_rank or not_posdef" 2>&1 | tail -3
python tools/c1_breakdown.py 2>&1 | head -3

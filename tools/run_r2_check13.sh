python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_reference_vectors.py -m gpu -q 2>&1 | tail -3
python tools/microbench_r2.py 20000 2>&1 | tail -1

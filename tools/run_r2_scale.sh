python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_reference_vectors.py tests/test_gpu_bart.py -m gpu -q -x 2>&1 | tail -2
python tools/time_step.py 20000 4 2>&1 | tail -1
python tools/time_step.py 10000 4 2>&1 | tail -1

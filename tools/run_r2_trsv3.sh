python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "vector_solves" -v 2>&1 | grep -E "PASS|FAIL|assert np|AssertionError" | head -40

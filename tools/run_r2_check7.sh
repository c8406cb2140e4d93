python -m pytest tests/test_gpu_bart.py tests/test_reference_vectors.py tests/test_gpu_api.py -m gpu -q -k "bart or BART" 2>&1 | tail -6
python tools/microbench_r2.py 20000 2>&1 | tail -1

python -m pytest tests/test_gpu_bart.py tests/test_gpu_api.py -m gpu -q -k "multistart or bayestree or c4_size" 2>&1 | tail -40
python tools/bench_configs.py 2>&1 | tail -3

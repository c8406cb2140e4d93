for cb in 0 32 64 1000; do echo "CHAIN=$cb: $(LGP_CHAIN_BLOCKS=$cb python tools/time_chol.py 20000,10000,5000,2048,1024 2>&1 | tail -1)"; done
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_reference_vectors.py -m gpu -q -x 2>&1 | tail -2

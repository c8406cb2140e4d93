for v in 1 31; do echo "LEAF=$v: $(LGP_LEAF=$v python tools/time_chol.py 20000,10000,4096,2048 2>&1 | tail -1)"; done
for v in 1 31; do echo "LEAF=$v: $(LGP_LEAF=$v python tools/time_chol.py 20000,10000,4096,2048 2>&1 | tail -1)"; done
python tools/trace_chol.py 20000 2>&1 | tail -45
python -m pytest tests/test_gpu_kernels.py tests/test_reference_vectors.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -3

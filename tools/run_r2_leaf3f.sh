for sb in 0 16 24 32 40; do echo "SERIAL=$sb: $(LGP_SERIAL_BLOCKS=$sb python tools/time_chol.py 20000,10000,4096,2048,1024 2>&1 | tail -1)"; done
echo "SERIAL=32 wide>16: $(LGP_SERIAL_BLOCKS=32 LGP_SERIAL_WIDE=16 python tools/time_chol.py 20000,10000,4096,2048,1024 2>&1 | tail -1)"
echo "SERIAL=48 wide>24: $(LGP_SERIAL_BLOCKS=48 LGP_SERIAL_WIDE=24 python tools/time_chol.py 20000,10000,4096,2048,1024 2>&1 | tail -1)"
echo "SERIAL=24 nosmall: $(LGP_GEMM_SMALL=0 python tools/time_chol.py 20000,10000,4096,2048,1024 2>&1 | tail -1)"
python -m pytest tests/test_gpu_kernels.py tests/test_reference_vectors.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -3

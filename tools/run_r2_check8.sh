for tb in "0,0" "48,20" "64,24" "32,12" "48,0" "80,32" "24,8"; do echo "TAIL=$tb: $(LGP_TAIL_BLOCKS=$tb python tools/time_chol.py 20000,10000,4096 2>&1 | tail -1)"; done
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "chol" 2>&1 | tail -3

for tb in "96,64,32" "96,48,32" "96,48,24" "96,40,24" "96,64,24" "96,32,16" "96,80,48" "96,56,40"; do echo "TAIL=$tb: $(LGP_TAIL_BLOCKS=$tb python tools/time_chol.py 10000,5000,2048 2>&1 | tail -1)"; done
python tools/time_solve.py 1000,4096,20000 2>&1 | tail -3
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "solves or chol" 2>&1 | tail -2

for rep in 1 2 3; do for tb in "1000000,64,24" "64,64,24" "80,48,20" "96,64,24" "100,64,24"; do echo "rep$rep TAIL=$tb: $(LGP_TAIL_BLOCKS=$tb python tools/time_chol.py 20000 2>&1 | tail -1)"; done; done
python -m pytest tests/test_gpu_api.py -m gpu -q -k "two_devices" 2>&1 | tail -2

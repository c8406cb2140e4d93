for tb in "1000000,64,24" "96,64,24" "112,64,24" "128,64,24" "80,48,20"; do echo "TAIL=$tb: $(LGP_TAIL_BLOCKS=$tb python tools/time_chol.py 20000,10000 2>&1 | tail -1)"; done

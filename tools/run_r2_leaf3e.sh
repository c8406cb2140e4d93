for tb in "96,64,24" "96,64,32" "96,64,40" "96,64,48" "96,64,64" "96,80,80" "96,96,96"; do echo "TAIL=$tb: $(LGP_TAIL_BLOCKS=$tb python tools/time_chol.py 20000,10000,4096 2>&1 | tail -1)"; done

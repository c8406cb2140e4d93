export LGP_SERIAL_BLOCKS=0
for tb in "96,64,24" "96,64,32" "96,64,40" "96,48,32" "96,80,32" "96,80,40" "128,64,32" "72,64,32"; do echo "TAIL=$tb: $(LGP_TAIL_BLOCKS=$tb python tools/time_chol.py 20000,10000,4096 2>&1 | tail -1)"; done
LGP_TAIL_BLOCKS=96,64,32 python tools/trace_chol.py 20000 2>&1 | tail -60 | head -50

for fb in 8 4 2 1; do echo "FIRST=$fb: $(LGP_FIRST_BLOCKS=$fb python tools/time_chol.py 20000,10000 2>&1 | tail -1)"; done
for fb in 8 4 2 1; do echo "FIRST=$fb: $(LGP_FIRST_BLOCKS=$fb python tools/time_chol.py 20000,10000 2>&1 | tail -1)"; done
LGP_FIRST_BLOCKS=2 python tools/trace_chol.py 20000 2>&1 | grep -A12 "n = 20000" | tail -13

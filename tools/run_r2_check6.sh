python tools/microbench_r2.py 20000 2>&1 | tail -2
LGP_GRAM_WAVES=0 python tools/microbench_r2.py 20000 2>&1 | tail -2
LGP_GRAM_WAVES=2 python tools/microbench_r2.py 20000 2>&1 | tail -2

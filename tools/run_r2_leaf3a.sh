python tools/bench_leaf.py 2>&1 | tail -20
for v in 1 3; do echo "LEAF=$v: $(LGP_LEAF=$v python tools/time_chol.py 20000 2>&1 | tail -1)"; done
for v in 1 3; do echo "LEAF=$v: $(LGP_LEAF=$v python tools/time_chol.py 10000 2>&1 | tail -1)"; done
for v in 1 3; do echo "LEAF=$v: $(LGP_LEAF=$v python tools/time_chol.py 4096 2>&1 | tail -1)"; done
LGP_LEAF=3 python -m pytest tests/test_gpu_kernels.py tests/test_reference_vectors.py -m gpu -q -x -k "chol" 2>&1 | tail -3

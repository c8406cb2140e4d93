python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5
for st in dense lower; do python tools/dist_check.py --size 40960 --tile 1024 --storage $st 2>&1 | tail -2 | cut -c1-900; done

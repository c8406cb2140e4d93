python -m pytest tests/ -m gpu -q 2>&1 | tail -4
python tools/time_solve.py 1000,4096,20000 2>&1 | tail -3

for i in 1 2 3; do python -m pytest tests/test_gpu_api.py -m gpu -q -k "config3_empbayes" 2>&1 | grep -E "assert|Error|passed|failed" | head -8; done
python -m pytest tests/ -m gpu -q 2>&1 | tail -8

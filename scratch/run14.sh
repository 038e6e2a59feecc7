python -m pytest tests -q -m gpu -x -k "random_sweep" 2>&1 | tail -15

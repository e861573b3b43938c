python -m pytest tests/test_layout_matrix.py -m gpu -q -x 2>&1 | tail -5

python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/single_chain.py 2>&1 | tail -3

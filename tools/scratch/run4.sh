python tools/dev_exp.py - 2>&1 | tail -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -5

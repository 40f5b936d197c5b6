python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/train_bench.py --batch 64 2>&1 | tail -1 | cut -c1-330
python tools/train_bench.py --batch 8 2>&1 | tail -1 | cut -c1-330

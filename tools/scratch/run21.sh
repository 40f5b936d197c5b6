python -m pytest tests/test_routing_gpu.py -x -q -m gpu -k "mixed_precision" 2>&1 | tail -5
for b in 64 8; do
python tools/train_bench.py --uhat bf16 --batch $b 2>&1 | tail -1 | cut -c80-260
python tools/train_bench.py --uhat f16 --bwd-uhat bf16 --batch $b 2>&1 | tail -1 | cut -c80-260
done

for c in 0 1; do
echo "== C16=$c"
SRF_BWD_C16=$c python tools/dev_bwd_kernel_times.py 8 375 bf16 2>&1 | grep "route_layer_bwd\|layer bwd"
SRF_BWD_C16=$c python tools/train_bench.py --uhat f16 --bwd-uhat bf16 --batch 8 2>&1 | tail -1 | cut -c150-260
done
SRF_BWD_C16=1 python -m pytest tests/test_routing_gpu.py tests/test_training_gpu.py -x -q -m gpu -k "backward or bwd or train or grad" 2>&1 | tail -2

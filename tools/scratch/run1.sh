set -x
python -m pytest tests/test_routing_gpu.py -x -q -m gpu -k "fused or full_size or bench_mode" 2>&1 | tail -5
python tools/dev_exp.py - SRF_FUSED_CAPSTAGE=0 2>&1 | tail -20

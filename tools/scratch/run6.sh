python -m pytest tests/test_routing_gpu.py -x -q -m gpu -k "fused or full_size or bench_mode" 2>&1 | tail -3
echo "== new"; python tools/dev_exp.py - 2>&1 | tail -3
echo "== base"; SRF_B200_LIB=tools/scratch/lib_base.so python tools/dev_exp.py - 2>&1 | tail -3
echo "== new"; python tools/dev_exp.py - 2>&1 | tail -3

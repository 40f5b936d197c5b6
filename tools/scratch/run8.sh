python -m pytest tests/test_routing_gpu.py -x -q -m gpu -k "fused or full_size or bench_mode" 2>&1 | tail -3
for v in ffma2 new ffma2 new; do
echo "== $v"; if [ $v = new ]; then python tools/dev_exp.py - 2>&1 | tail -3; else SRF_B200_LIB=tools/scratch/lib_$v.so python tools/dev_exp.py - 2>&1 | tail -3; fi
done

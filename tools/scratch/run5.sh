for m in bf16 f16; do
python tools/dev_bwd_kernel_times.py 64 375 $m 2>&1 | tail -8
python tools/dev_bwd_kernel_times.py 8 375 $m 2>&1 | tail -8
python tools/train_bench.py --uhat $m --batch 64 2>&1 | tail -1 | cut -c1-400
python tools/train_bench.py --uhat $m --batch 8 2>&1 | tail -1 | cut -c1-400
done

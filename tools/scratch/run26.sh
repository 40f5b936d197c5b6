python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 300 gpurun_out/r2d_bench_n1.err
cut -c1-250 gpurun_out/r2d_bench_n1.json

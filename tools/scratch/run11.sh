python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 600 gpurun_out/r2c_bench_n1.err
cat gpurun_out/r2c_bench_n1.json | cut -c1-1500

N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err
tail -c 400 gpurun_out/r2d_bench_n$N.err; cut -c1-200 gpurun_out/r2d_bench_n$N.json

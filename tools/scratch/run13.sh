set -x
python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/r2c_plain.json 2> gpurun_out/r2c_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches_cfg3_f16.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/r2c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:route_fused -s 3 -c 1 -o gpurun_out/r2c_fused_f16 -f python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/r2c_ncu2.log 2>&1
SRF_B200_LIB=tools/scratch/lib_timers.so python tools/dev_fused_timers.py cfg3 f16 > gpurun_out/r2c_timers.txt 2>&1
tail -22 gpurun_out/r2c_timers.txt

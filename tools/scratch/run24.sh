python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 300 gpurun_out/r2d_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['strong']['ms_per_step'], d['train_cfg4']['ms_per_step'], d['train_cfg4']['uhat'])
print({k:v.get('ms_per_step') for k,v in d['also'].items()})
PY
ncu --set full --clock-control none --import-source on -k regex:route_fused -s 3 -c 1 -o gpurun_out/r2d_fused_f16 -f python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/r2d_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2d_launches_cfg3_f16.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/r2d_ncu1.log 2>&1

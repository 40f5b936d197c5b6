for i in 1 2 3; do
for na in 1 0; do
if [ $na = 1 ]; then export SRF_FUSED_NO_ALIGN=1; else unset SRF_FUSED_NO_ALIGN; fi
echo -n "NO_ALIGN=$na  "; python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['e2e']['value']))"
done; done

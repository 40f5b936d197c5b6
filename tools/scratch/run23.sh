for w in 0 1; do
echo "== L2_WINDOW=$w"
if [ $w = 1 ]; then export SRF_L2_WINDOW=1; fi
python tools/dev_exp.py - 2>&1 | tail -3 | head -1 | cut -c1-40
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:route_fused -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline 2>&1 | grep -E "dram__|lts__|gpu__time"
done

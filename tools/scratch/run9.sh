for v in cur own2 sleep200 sleep20 cur own2 sleep200 sleep20; do
echo "== $v"; SRF_B200_LIB=tools/scratch/lib_$v.so python tools/dev_exp.py - 2>&1 | tail -3 | head -2
done

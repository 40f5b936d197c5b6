for v in cur v1 v2 v3 cur v1 v2 v3; do
echo "== $v"; SRF_B200_LIB=tools/scratch/lib_$v.so python tools/dev_exp.py - 2>&1 | tail -3 | head -1
done

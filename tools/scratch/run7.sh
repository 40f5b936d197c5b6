for v in base ffma2 base ffma2; do
echo "== $v"; SRF_B200_LIB=tools/scratch/lib_$v.so python tools/dev_exp.py - 2>&1 | tail -3
done

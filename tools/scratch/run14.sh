for v in cur iss cur iss; do
echo "== $v"; SRF_B200_LIB=tools/scratch/lib_$v.so python tools/dev_exp.py - 2>&1 | tail -3
done

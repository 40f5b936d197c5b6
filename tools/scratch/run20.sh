for v in w24 w32; do
echo "== $v"
SRF_B200_LIB=gpurun_variants/lib_$v.so python tools/dev_bwd_kernel_times.py 64 375 bf16 2>&1 | grep "dwdx\|layer bwd"
SRF_B200_LIB=gpurun_variants/lib_$v.so python tools/dev_bwd_kernel_times.py 8 375 bf16 2>&1 | grep "dwdx\|layer bwd"
done

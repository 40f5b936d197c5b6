SRF_B200_LIB=tools/scratch/lib_timers.so python tools/dev_fused_timers.py cfg3 f16 2>&1 | tail -25
SRF_B200_LIB=tools/scratch/lib_xst4.so python tools/dev_exp.py - 2>&1 | tail

#!/bin/bash
# c4 thread kernel with the L2 prefetch; ncu --set full (with source) of the n = 8192 register-table kernel
mkdir -p gpurun_out
python tools/profile_kernels.py --which decode_c4,decode_warp > gpurun_out/r2b_c4.log 2>&1
cat gpurun_out/r2b_c4.log
python tools/profile_kernels.py --which decode_block --c8k-codewords 8000 > gpurun_out/r2b_c8k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_regular_rt -s 2 -c 1 -o gpurun_out/r2b_rt \
    python tools/profile_kernels.py --which decode_block --c8k-codewords 8000 > gpurun_out/r2b_ncu_rt.log 2>&1
cat gpurun_out/r2b_c8k_plain.log

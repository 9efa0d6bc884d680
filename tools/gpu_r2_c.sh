#!/bin/bash
# where do 5-iteration launches spend their fixed cost?  ncu --set full with source of both C4 kernels at 5 iterations
mkdir -p gpurun_out
python tools/profile_kernels.py --which decode_c4,decode_warp --c4-codewords 1000000 > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_c4_thread -s 2 -c 1 -o gpurun_out/r2c_c4_5it \
    python tools/profile_kernels.py --which decode_c4 --c4-codewords 1000000 > gpurun_out/r2c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_warp -s 2 -c 1 -o gpurun_out/r2c_warp_5it \
    python tools/profile_kernels.py --which decode_warp --c4-codewords 1000000 > gpurun_out/r2c_ncu2.log 2>&1
cat gpurun_out/r2c_plain.log

#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_run.sh r2ab new,r100,r103,r108 tools/profile_kernels.py --which decode_c4 2>&1 | grep -v "^$" > gpurun_out/r2ab_times.txt
cat gpurun_out/r2ab_times.txt

#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_build.py tests/test_gpu_blocks.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r2w_tests.txt
cat gpurun_out/r2w_tests.txt
for cfg in "warp 0 2 5 1" "warp 0 6 5 1" "warp 1 2 5 1" "warp 1 6 5 1" "warp 1 8 5 1" "warp 2 2 5 1" "auto 1 6 5 1"; do
  python tools/warp_one.py $cfg 2>&1 | grep "dB iters"
done > gpurun_out/r2w_times.txt
cat gpurun_out/r2w_times.txt

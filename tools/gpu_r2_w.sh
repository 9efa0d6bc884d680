#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_build.py tests/test_gpu_blocks.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2w_tests.txt; cat gpurun_out/r2w_tests.txt
for cfg in "warp 0 2 5 1" "warp 0 6 5 1" "warp 0 2 50 0"; do python tools/warp_one.py $cfg 2>&1 | grep "dB iters"; done

#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2ab_tests.txt; cat gpurun_out/r2ab_tests.txt
for r in 1 2; do python tools/profile_kernels.py --which decode_c4,decode_block --c8k-codewords 8000 2>&1 | grep "decode_"; done > gpurun_out/r2ab_times.txt
cat gpurun_out/r2ab_times.txt

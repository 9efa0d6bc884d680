#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_build.py tests/test_gpu_blocks.py -m gpu -x -q -k "enc or c8k or headless" 2>&1 | tail -3 ) > gpurun_out/r2e_tests.txt; cat gpurun_out/r2e_tests.txt
python tools/enc_sweep.py --frames 100,500,1000,2000,3000 --configs "ring=60" --reps 20 > gpurun_out/r2e_small.txt 2>&1; cat gpurun_out/r2e_small.txt

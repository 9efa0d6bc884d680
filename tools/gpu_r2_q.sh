#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocks.py -m gpu -x -q -k "host_path or ragged or headless or pool" 2>&1 | tail -3 ) > gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_tests.txt
MODES="-1 1" bash tools/gpu_r2_n.sh 1

#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "enc or host_path" 2>&1 | tail -3 ) > gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_tests.txt
python tools/enc_sweep.py --frames 1000,2000,2048,3000 --configs "ring=60" --reps 20 2>&1 | grep ring

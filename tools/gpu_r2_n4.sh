#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "enc" 2>&1 | tail -2 ) > gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_tests.txt
python tools/enc_sweep.py --frames 500,1000,2000,3000 --configs "ring=60" --reps 20 2>&1 | grep ring; 
MODES=-1 bash tools/gpu_r2_n.sh 4

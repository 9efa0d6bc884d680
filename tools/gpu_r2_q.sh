#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "enc" 2>&1 | tail -3 ) > gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_tests.txt
python tools/enc_sweep.py --frames 20000,50000,100000,200000,400000 --configs "ring=60;ring=60,tile=868;ring=60" --reps 5 2>&1 | grep ring

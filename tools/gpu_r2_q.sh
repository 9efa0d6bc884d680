#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "enc" 2>&1 | tail -2 ) > gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_tests.txt
python __graft_entry__.py smoke 2>&1 | grep "smoke ok"
python tools/enc_sweep.py --frames 2048,20000,200000 --configs "ring=60" 2>&1 | grep ring

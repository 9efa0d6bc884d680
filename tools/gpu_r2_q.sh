#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "adaptive or families or golden" 2>&1 | tail -2 ) > gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_tests.txt
python __graft_entry__.py smoke 2>&1 | grep "smoke ok"

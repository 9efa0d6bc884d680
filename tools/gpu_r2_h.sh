#!/bin/bash
# round-2 re-entry check on the GPU box: full GPU test suite + default bench line at HEAD
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2h_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_tests.log 2>&1
tail -3 gpurun_out/r2h_tests.log
( time python bench.py ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -c 400 gpurun_out/r2h_bench.json
python tools/profile_kernels.py --which encode_generic --c8k-codewords 200000 2>&1 | grep -i "encode" > gpurun_out/r2h_enc200k.log
python tools/profile_kernels.py --which encode_generic --c8k-codewords 400000 2>&1 | grep -i "encode" >> gpurun_out/r2h_enc200k.log
cat gpurun_out/r2h_enc200k.log

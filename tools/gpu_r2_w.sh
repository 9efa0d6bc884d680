#!/bin/bash
mkdir -p gpurun_out
for m in 1 2 4 16 64; do for cfg in "warp 1 6 5 1" "warp 0 2 5 1"; do
  LDPC535_WARP_GRID_MULT=$m python tools/warp_one.py $cfg 10000000 2>&1 | grep "dB iters" | sed "s/^/mult $m: /"
done; done > gpurun_out/r2w_gridmult.txt
cat gpurun_out/r2w_gridmult.txt

#!/bin/bash
mkdir -p gpurun_out
for lib in new mb6 mb5; do
for cfg in "warp 0 2 5 1" "warp 0 6 5 1" "warp 1 2 5 1" "warp 1 6 5 1" "warp 1 8 5 1" "warp 1 2 50 0"; do
  if [ $lib = new ]; then python tools/warp_one.py $cfg 2>&1 | grep "dB iters" | sed "s/^/$lib /"; else LDPC535_LIB=$PWD/tools/ab/libldpc535_$lib.so python tools/warp_one.py $cfg 2>&1 | grep "dB iters" | sed "s/^/$lib /"; fi
done; done > gpurun_out/r2w_times.txt
cat gpurun_out/r2w_times.txt

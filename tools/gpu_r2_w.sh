#!/bin/bash
mkdir -p gpurun_out
for cfg in "warp 0 2 5 1"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_warp_kernel -s 1 -c 1 -o gpurun_out/r2w_$tag -f python tools/warp_one.py $cfg > gpurun_out/r2w_ncu_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2w_$tag.ncu-rep > gpurun_out/r2w_$tag.txt 2>&1
done

#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_c4_thread -s 1 -c 1 -o gpurun_out/r2p_c4_5it -f python tools/warp_one.py c4-thread 1 2 5 1 > gpurun_out/r2p_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2p_c4_5it.ncu-rep > gpurun_out/r2p_c4_5it.txt 2>&1

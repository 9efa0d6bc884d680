#!/bin/bash
mkdir -p gpurun_out
for cfg in "warp 1 6 5 1" "warp 0 2 5 1"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_warp_kernel -s 1 -c 1 -o gpurun_out/r2w_$tag -f python tools/warp_one.py $cfg > gpurun_out/r2w_ncu_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2w_$tag.ncu-rep > gpurun_out/r2w_$tag.txt 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:encode_m4r_(fr_)?kernel' -s 1 -c 1 -o gpurun_out/r2j_enc_607 -f \
    python tools/enc_sweep.py --frames 200000 --reps 1 --configs "ring=60,tpf=7" > gpurun_out/r2j_ncu_607.log 2>&1
python tools/ncu_summary.py gpurun_out/r2j_enc_607.ncu-rep > gpurun_out/r2j_enc_607.txt 2>&1

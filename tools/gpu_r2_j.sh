#!/bin/bash
# ncu --set full of the look-up encoder variants (200 000 frames)
mkdir -p gpurun_out
for r in 42 60; do
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:encode_m4r_(fr_)?kernel' -s 1 -c 1 -o gpurun_out/r2j_enc_$r -f \
    python tools/enc_sweep.py --frames 200000 --reps 1 --configs "ring=$r,tile=992" > gpurun_out/r2j_ncu_$r.log 2>&1
  tail -2 gpurun_out/r2j_ncu_$r.log
  python tools/ncu_summary.py gpurun_out/r2j_enc_$r.ncu-rep > gpurun_out/r2j_enc_$r.txt 2>&1
done

#!/bin/bash
# look-up encoder: parity test over the variants, barrier kernel vs free-running kernel, ncu of the default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encoder" 2>&1 | tail -5 > gpurun_out/r2i_enc_tests.txt
cat gpurun_out/r2i_enc_tests.txt
timeout 300 python tools/enc_sweep.py --frames 3072,8000,20000,200000,400000 --configs "ring=42,tile=1024;ring=60,tpf=7;ring=60,tpf=7;ring=42,tile=1024" > gpurun_out/r2i_enc_sweep.txt 2>&1
cat gpurun_out/r2i_enc_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:encode_m4r_(fr_)?kernel' -s 1 -c 1 -o gpurun_out/r2j_enc_607 -f \
    python tools/enc_sweep.py --frames 200000 --reps 1 --configs "ring=60,tpf=7" > gpurun_out/r2j_ncu_607.log 2>&1
python tools/ncu_summary.py gpurun_out/r2j_enc_607.ncu-rep > gpurun_out/r2j_enc_607.txt 2>&1

#!/bin/bash
# round-2 baseline on the GPU box: full GPU test suite, default bench line, launch list, ncu full of the headline kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2g_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2g_tests.log 2>&1
tail -3 gpurun_out/r2g_tests.log
( time python bench.py ) > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -c 600 gpurun_out/r2g_bench.json
python bench.py --no-cpu --no-extras --steps 2 > gpurun_out/r2g_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py --no-cpu --no-extras --steps 2 > gpurun_out/r2g_ncu_launch.log 2>&1
python tools/profile_kernels.py --which decode_c4 > gpurun_out/r2g_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_c4_thread -c 1 -o gpurun_out/r2g_c4 \
    python tools/profile_kernels.py --which decode_c4 > gpurun_out/r2g_ncu_c4.log 2>&1
cat gpurun_out/r2g_prof_plain.log

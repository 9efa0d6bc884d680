#!/bin/bash
# round-2 final evidence on the GPU box: full GPU test suite, default bench line, reference arm, launch list,
# ncu --set full of the headline kernel, the n = 8192 decoder and the look-up encoder
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2z_smi.txt 2>&1
( time python __graft_entry__.py smoke ) > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_tests.log 2>&1
tail -3 gpurun_out/r2z_tests.log
( time python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -c 300 gpurun_out/r2z_bench.json; tail -3 gpurun_out/r2z_bench.err
( time python bench.py --impl reference ) > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
tail -c 300 gpurun_out/r2z_bench_ref.json
python bench.py --no-cpu --no-extras --steps 2 > gpurun_out/r2z_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv \
    python bench.py --no-cpu --no-extras --steps 2 > gpurun_out/r2z_ncu_launch.log 2>&1
python tools/profile_kernels.py --which decode_c4 > gpurun_out/r2z_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_c4_thread -c 1 -o gpurun_out/r2z_c4 -f \
    python tools/profile_kernels.py --which decode_c4 > gpurun_out/r2z_ncu_c4.log 2>&1
python tools/ncu_summary.py gpurun_out/r2z_c4.ncu-rep > gpurun_out/r2z_c4.txt 2>&1
python tools/profile_kernels.py --which decode_block --c8k-codewords 8000 > gpurun_out/r2z_prof_c8k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_regular_rt -s 1 -c 1 -o gpurun_out/r2z_rt -f \
    python tools/profile_kernels.py --which decode_block --c8k-codewords 8000 > gpurun_out/r2z_ncu_rt.log 2>&1
python tools/ncu_summary.py gpurun_out/r2z_rt.ncu-rep > gpurun_out/r2z_rt.txt 2>&1
cat gpurun_out/r2z_prof_plain.log gpurun_out/r2z_prof_c8k.log | grep -v "Exception\|Traceback\|File \|AttributeError"
python tools/refill_sweep.py 4000000 2>&1 | grep "Eb/N0" > gpurun_out/r2z_early_stop_families.txt; cat gpurun_out/r2z_early_stop_families.txt
python tools/enc_sweep.py --frames 500,2000,3072,8000,20000,200000,400000 --configs "ring=60" > gpurun_out/r2z_enc_sweep.txt 2>&1; cat gpurun_out/r2z_enc_sweep.txt
for cfg in "warp 0 2 5 1" "warp 0 6 5 1" "warp 2 2 5 1"; do python tools/warp_one.py $cfg 2>&1 | grep "dB iters"; done > gpurun_out/r2z_warp_times.txt; cat gpurun_out/r2z_warp_times.txt

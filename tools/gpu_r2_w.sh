#!/bin/bash
mkdir -p gpurun_out
python tools/refill_sweep.py 4000000 2>&1 | grep "Eb/N0" > gpurun_out/r2n_early_stop_families.txt
cat gpurun_out/r2n_early_stop_families.txt

#!/bin/bash
mkdir -p gpurun_out
python tools/run_configs.py 2>&1 | grep -v "SYNC\|^Method:\|^BMP Header\|^File written\|Exception ignored\|Traceback\|  File \|AttributeError" > gpurun_out/r2z_configs.md
tail -40 gpurun_out/r2z_configs.md

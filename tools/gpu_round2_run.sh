python -m pytest tests -m gpu -x -q -k "mandril or ber_fer or refill or reference_image" 2>&1 | tail -8 > gpurun_out/r2_t5.log
python tools/refill_sweep.py 2>&1 | grep -v "Exception ignored\|Traceback\|File \|AttributeError" > gpurun_out/r2_refill_sweep3.log
python tools/run_configs.py 2 2>&1 | grep -v "SYNC" > gpurun_out/r2_config2.md
tail -4 gpurun_out/r2_t5.log; cat gpurun_out/r2_refill_sweep3.log; grep "^|" gpurun_out/r2_config2.md

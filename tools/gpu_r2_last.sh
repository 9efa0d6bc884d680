#!/bin/bash
# last check of the round on the committed tree: full GPU suite, smoke, default bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2last_tests.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2last_tests.log | tail -2
( time python __graft_entry__.py smoke ) > gpurun_out/r2last_smoke.log 2>&1; grep -E "smoke ok|Error" gpurun_out/r2last_smoke.log
( time python bench.py ) > gpurun_out/r2last_bench.json 2> gpurun_out/r2last_bench.err; tail -3 gpurun_out/r2last_bench.err
python - <<PY
import json
for l in open('gpurun_out/r2last_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.3f e2e %.3f frac %.3f launches %d'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['gpu_launches']))
PY

#!/bin/bash
# last check of the round on the committed tree: full GPU suite and smoke
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2last_tests.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2last_tests.log | tail -2
python __graft_entry__.py smoke 2>&1 | grep -E "smoke ok|Error"

#!/bin/bash
# A/B on one box: libraries under tools/ab/ (builds of an earlier commit or of an experiment) against the working tree's
# usage: tools/ab_run.sh <tag> <lib names without libldpc535_ prefix, comma separated, "new" = working tree> <python script and args...>
tag=$1; libs=$2; shift 2
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader > gpurun_out/${tag}_smi.txt
for round in 1 2; do
  for l in ${libs//,/ }; do
    echo "== $l (round $round)"
    if [ "$l" = new ]; then python "$@" 2>&1 | grep -v "Exception ignored\|Traceback\|File \|AttributeError"
    else LDPC535_LIB=$PWD/tools/ab/libldpc535_$l.so python "$@" 2>&1 | grep -v "Exception ignored\|Traceback\|File \|AttributeError"; fi
  done
done

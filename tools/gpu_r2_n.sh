#!/bin/bash
# multi-GPU bench line (one rank per GPU), as the driver launches it; then the host path forced to raw / packed for comparison
N=$1
mkdir -p gpurun_out
for mode in ${MODES:--1}; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-extras --no-cpu --pack-pinned $mode ) > gpurun_out/r2_bench_n${N}_pack$mode.json 2> gpurun_out/r2_bench_n${N}_pack$mode.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench_n${N}_pack$mode.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('N=$N mode $mode value %.3f e2e %.3f Gbit/s %.1f ms | pcie GB/step %.2f | ceiling %.1f GB/s frac %.2f | %s'%(d['value'],e['value'],e['ms_per_step'],e['pcie_h2d_bytes_per_step']/1e9,e['h2d_ceiling_gbs'],e['frac_of_h2d_ceiling'],e['host_path'][:200]))
PY
done

#!/bin/bash
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2_n2_full.json 2> gpurun_out/r2_n2_full.err
tail -4 gpurun_out/r2_n2_full.err
python - <<PY
import json
for l in open('gpurun_out/r2_n2_full.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n',d['n_gpus'],'value %.3f e2e %.3f'%(d['value'],d['e2e']['value']), 'configs' in d, 'cpu_baseline' in d)
PY
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r2_n2_ref.json 2> gpurun_out/r2_n2_ref.err
tail -c 300 gpurun_out/r2_n2_ref.json; tail -3 gpurun_out/r2_n2_ref.err

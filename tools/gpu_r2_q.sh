#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02q_n2.json 2>> $L
echo "n2 exit $?" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02q_ref2.json 2>> $L
echo "ref2 exit $?" >> $L
python - <<PY >> $L
import json
d=json.loads(open('gpurun_out/r02q_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['ranks'], list(d.keys()))
print(d.get('weak_scaling'))
print(open('gpurun_out/r02q_ref2.json').read()[:400])
PY
grep -v "^$" $L | grep -v "OMP_NUM\|\*\*\*" | tail -20

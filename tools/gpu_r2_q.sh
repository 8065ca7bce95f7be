#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02q_n1.json 2>> $L
echo "n1 exit $?" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extras --no-cpu-baseline > gpurun_out/r02q_n2.json 2>> $L
echo "n2 exit $?" >> $L
python - <<PY >> $L
import json
for f in ('n1','n2'):
    d=json.loads(open('gpurun_out/r02q_%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['ranks'])
PY
grep -v "^$" $L | tail -30

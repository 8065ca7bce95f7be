#!/bin/bash
# same box: N=1 then N=8 (strong scaling, one 4096-crop batch), short lines without the extras
set -u
mkdir -p gpurun_out
L=gpurun_out/r02n8.log
: > $L
true
echo "n1 exit $?" >> $L
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --no-cpu-baseline --no-extras > gpurun_out/r02n8_n8.json 2>> $L
echo "n8 exit $?" >> $L
python - <<PY >> $L
import json
for f in ('n8',):
    d=json.loads(open('gpurun_out/r02n8_%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['ranks'], d['clocks'])
PY
grep -v "^$" $L | grep -v "OMP_NUM\|\*\*\*" | tail -20

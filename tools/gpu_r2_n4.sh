#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02n4.log
: > $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --no-cpu-baseline --no-extras > gpurun_out/r02n4_n4.json 2>> $L
echo "n4 exit $?" >> $L
python - <<PY >> $L
import json
d=json.loads(open('gpurun_out/r02n4_n4.json').read().strip().splitlines()[-1])
print('n4', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['ranks'], d['clocks'])
PY
grep -v "^$" $L | grep -v "OMP_NUM\|\*\*\*" | tail -6

#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
for g in auto on auto on; do
timeout 600 python bench.py --graph $g --no-extras --no-cpu-baseline > gpurun_out/r02q_$g.json 2>> $L
python - <<PY >> $L
import json
d=json.loads(open('gpurun_out/r02q_$g.json').read().strip().splitlines()[-1])
print('$g', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'].get('power_w_max'), d['roofline']['achieved'], d['roofline'].get('timed'))
PY
done
grep -v "^$" $L | tail -70

#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 900 python -m pytest tests/test_gpu_bf16_layers.py tests/test_gpu_reference.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "matched oracle|passed|failed|Error" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke exit $?" >> $L
for b in 512 4096; do timeout 300 python tools/stage_breakdown.py a $b bf16 >> $L 2>&1; done
grep -v "^$" $L | tail -70

#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "tps" 2>&1 | tail -3 >> $L
for b in 512 4096; do timeout 300 python tools/stage_breakdown.py a $b bf16 2>&1 | grep -E "first|sum|maxpool|flat|tps" >> $L; done
grep -v "^$" $L | tail -70

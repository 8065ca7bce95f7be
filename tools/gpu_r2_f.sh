#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02f.log
: > $L
timeout 300 python tools/tune_seq.py 512 4096 >> $L 2>&1
for w in 4 2; do echo "=== MSTR_DECODE_W=$w" >> $L; MSTR_DECODE_W=$w timeout 300 python tools/tune_seq.py 4096 2>&1 | grep "attn_decode\|phase trace" >> $L; done
echo "tune exit $?" >> $L
tail -40 $L

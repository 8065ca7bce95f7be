#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02d.log
: > $L
timeout 300 python tools/tune_seq.py 256 512 4096 >> $L 2>&1
echo "tune exit $?" >> $L
timeout 300 python tools/graph_check.py 256 512 4096 >> $L 2>&1
echo "graph exit $?" >> $L
tail -30 $L

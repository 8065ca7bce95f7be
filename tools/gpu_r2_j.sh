#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02j.log
: > $L
timeout 200 python tools/diag_flat.py 4096 >> $L 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_flat -s 2 -c 1 -o gpurun_out/r02_flat python tools/diag_flat.py 4096 >> gpurun_out/r02j_ncu.log 2>&1
echo "ncu exit $?" >> $L
timeout 300 python bench.py --batch 512 --steps 20 --warmup 3 --no-extras --no-cpu-baseline --graph off > gpurun_out/r02j_bench512.json 2>> $L
echo "bench512 exit $?" >> $L
tail -5 $L

#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s -k "fiducial" 2>&1 | grep -E "C' vs|passed|failed|Error|rel-to-max" >> $L
grep -v "^$" $L | tail -70

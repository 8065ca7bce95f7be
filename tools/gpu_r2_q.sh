#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "ragged" 2>&1 | tail -15 >> $L
grep -v "^$" $L | tail -70

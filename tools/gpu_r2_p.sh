#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02p.log
: > $L
timeout 900 python -m pytest tests/test_gpu_transformer.py -q -m gpu -p no:cacheprovider -s >> $L 2>&1
echo "tf tests exit $?" >> $L
grep -v "^$" $L | tail -60

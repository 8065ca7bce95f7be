#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02n.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "attn_decode" >> $L 2>&1
echo "decoder tests exit $?" >> $L
timeout 400 python tools/tune_seq.py 256 512 1024 2048 4096 >> $L 2>&1
echo "tune exit $?" >> $L
tail -60 $L

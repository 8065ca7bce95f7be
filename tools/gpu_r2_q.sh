#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "attn or decode" 2>&1 | tail -3 >> $L
timeout 300 python tools/tune_seq.py 256 512 1024 1664 2048 2>&1 | grep -E "attn_decode|phase trace|fine cell tile 0" >> $L
timeout 400 python tools/soak_decoder.py >> $L 2>&1
echo "soak exit $?" >> $L
grep -v "^$" $L | tail -40

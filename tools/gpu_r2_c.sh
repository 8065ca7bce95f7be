#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02c.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x -k "attn_decode_bf16" >> $L 2>&1
echo "decoder tests exit $?" >> $L
timeout 300 python tools/tune_seq.py 256 512 4096 >> $L 2>&1
echo "tune exit $?" >> $L
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference.py -m gpu -q -p no:cacheprovider -k "bf16 or drop_in" >> $L 2>&1
echo "model tests exit $?" >> $L
tail -40 $L

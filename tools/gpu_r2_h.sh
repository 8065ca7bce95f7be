#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02i.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -s -k "padded_flat or conv3x3_first or maxpool" >> $L 2>&1
echo "flat tests exit $?" >> $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x -k "attn_decode_bf16" >> $L 2>&1
echo "decoder tests exit $?" >> $L
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_bf16_layers.py -m gpu -q -p no:cacheprovider >> $L 2>&1
echo "model tests exit $?" >> $L
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_bench.json 2>> $L
echo "bench exit $?" >> $L
grep -v "^$" $L | tail -60

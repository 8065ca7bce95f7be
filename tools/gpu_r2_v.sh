#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02v.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "padded_flat" >> $L 2>&1
echo "flat tests exit $?" >> $L
timeout 600 python -m pytest tests/test_gpu_bf16_layers.py -q -m gpu -p no:cacheprovider >> $L 2>&1
echo "layer tests exit $?" >> $L
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02v_bench.json 2>> $L
echo "bench exit $?" >> $L
grep -v "^$" $L | tail -25

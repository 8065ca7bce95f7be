#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02k.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "padded_flat" >> $L 2>&1
echo "flat tests exit $?" >> $L
timeout 200 python tools/diag_flat.py 4096 >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_bf16_layers.py -m gpu -q -p no:cacheprovider >> $L 2>&1
echo "layer tests exit $?" >> $L
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02k_bench.json 2>> $L
echo "bench exit $?" >> $L
grep -v "^$" $L | tail -12

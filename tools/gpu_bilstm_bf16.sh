#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02x.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "bilstm" -x >> $L 2>&1
echo "bilstm tests exit $?" >> $L
timeout 900 python -m pytest tests/test_gpu_bf16_layers.py tests/test_gpu_model.py tests/test_gpu_reference.py -q -m gpu -p no:cacheprovider -x >> $L 2>&1
echo "model tests exit $?" >> $L
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02x_bench.json 2>> $L
echo "bench exit $?" >> $L
grep -v "^$" $L | tail -30

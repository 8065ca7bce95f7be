#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02o.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -x >> $L 2>&1
echo "tests exit $?" >> $L
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02o_bench.json 2>> $L
echo "bench exit $?" >> $L
timeout 300 python bench.py --batch 512 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02o_bench512.json 2>> $L
echo "bench512 exit $?" >> $L
tail -12 $L

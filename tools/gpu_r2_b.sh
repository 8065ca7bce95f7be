#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference.py tests/test_gpu_bf16_layers.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r02b_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02b_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench exit $?"
tail -5 gpurun_out/r02b_tests.log

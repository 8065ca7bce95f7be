#!/bin/bash
# round 2, call A: whole GPU suite (new parity tests included) + the 1-GPU bench line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/r02a_tests.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s >> gpurun_out/r02a_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench exit $?"
tail -5 gpurun_out/r02a_tests.log

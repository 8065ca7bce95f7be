#!/bin/bash
# round 2, call M: whole GPU suite as the driver runs it (-x) + smoke + default bench + reference arm
set -u
mkdir -p gpurun_out
L=gpurun_out/verify.log
: > $L
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider >> $L 2>&1
echo "pytest exit $?" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke exit $?" >> $L
timeout 900 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err
echo "bench exit $?" >> $L
timeout 600 python bench.py --impl reference > gpurun_out/verify_ref.json 2> gpurun_out/verify_ref.err
echo "ref exit $?" >> $L
tail -15 $L

#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 >> $L
echo "pytest exit ${PIPESTATUS[0]}" >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $L
echo "smoke exit ${PIPESTATUS[0]}" >> $L
grep -v "^$" $L | tail -12

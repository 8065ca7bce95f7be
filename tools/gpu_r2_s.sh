#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02s.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "bilstm or attn_decode" >> $L 2>&1
echo "tests exit $?" >> $L
timeout 300 python tools/tune_seq.py 4096 2>&1 | grep "bilstm\|attn_decode" >> $L
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:bilstm_tc -s 2 -c 1 -o gpurun_out/r02s_bilstm python tools/tune_seq.py 4096 > gpurun_out/r02s_ncu.log 2>&1
echo "ncu exit $?" >> $L
tail -8 $L

#!/bin/bash
# Final-state evidence of round 2: launch list of one bench step + --set full captures of the kernels changed late
# (each command has exited 0 without ncu before).
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$BENCH > gpurun_out/r02c_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
LPS=$(python -c "import json; print(json.loads([l for l in open('gpurun_out/r02c_plain.log') if l.startswith('{')][-1])['gpu_launches'])")
echo "launches per step: $LPS"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mstr:: -s $((3 * LPS)) -c $LPS --csv --log-file gpurun_out/r02c_launches.csv $BENCH > gpurun_out/r02c_ncu0.log 2>&1; echo "launch list $?"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k regex:bilstm_tc -s 6 -c 1 -o gpurun_out/r02c_bilstm_tc $BENCH > gpurun_out/r02c_ncu1.log 2>&1; echo "bilstm_tc $?"
$NCU -k regex:conv_gemm_bf16_2cta -s $((31 * 3 + 18)) -c 1 -o gpurun_out/r02c_conv $BENCH > gpurun_out/r02c_ncu2.log 2>&1; echo "conv $?"
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --kernel-table gpurun_out/r02c_kernel_table.json > gpurun_out/r02c_bench.log 2>&1; echo "kernel table $?"

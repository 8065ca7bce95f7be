#!/bin/bash
# round 2, last capture: launch list of one bench step of the FINAL library + --set full of the localisation head
# (the plain command has exited 0 before each ncu pass)
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$BENCH > gpurun_out/r02e_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
LPS=$(python -c "import json; print(json.loads([l for l in open('gpurun_out/r02e_plain.log') if l.startswith('{')][-1])['gpu_launches'])")
echo "launches per step: $LPS"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mstr:: -s $((3 * LPS)) -c $LPS --csv --log-file gpurun_out/r02e_launches.csv $BENCH > gpurun_out/r02e_ncu0.log 2>&1; echo "launch list $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:locnet_head -s 3 -c 1 -o gpurun_out/r02e_locnet_head $BENCH > gpurun_out/r02e_ncu1.log 2>&1; echo "locnet_head $?"
ls -la gpurun_out/r02e_*

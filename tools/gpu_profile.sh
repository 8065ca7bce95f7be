#!/bin/bash
# ncu evidence for bench.py (B200_PROFILING.md recipe): (1) the launch list of one timed step,
# (2) one --set full capture of the dominant kernel (512->512 3x3 tcgen05 implicit GEMM).
# Usage on the GPU box: bash tools/gpu_profile.sh <tag>     -> gpurun_out/<tag>_*.{csv,ncu-rep,log}
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log
# launches per step (LAUNCHES_PER_STEP is printed by the plain run's JSON: gpu_launches / steps)
LPS=$(python -c "import json,sys; print(json.loads(open('gpurun_out/${TAG}_plain.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per step: $LPS"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mstr:: \
    -s $((3 * LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_bf16 -s $((39 * 3 + 18)) -c 2 \
    -o gpurun_out/${TAG}_conv $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
# the other kernels of the path, one launch each in a steady-state step: TPS, both BiLSTM layers, decoder head/ctx/cell
ncu --set full --clock-control none --import-source on -k regex:"tps_rectify|bilstm_tc|attn_context|rnn_gemm" \
    -s $((3 * 82)) -c 7 -o gpurun_out/${TAG}_seq $CMD > gpurun_out/${TAG}_ncu_seq.log 2>&1
echo "ncu seq exit $?"
python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/${TAG}_kernel_table.json > gpurun_out/${TAG}_bench.log 2>&1
tail -1 gpurun_out/${TAG}_bench.log > gpurun_out/${TAG}_bench_line.json
ls -la gpurun_out/

#!/bin/bash
# ncu evidence for bench.py (B200_PROFILING.md recipe): (1) the launch list of one timed step, (2) one --set full capture
# of the dominant kernel (512->512 3x3 tcgen05 implicit GEMM), (3) --set full captures of the other kernels of the path.
# Every ncu command runs only after the same command has exited 0 without ncu.
# Usage on the GPU box: bash tools/gpu_profile.sh <tag>     -> gpurun_out/<tag>_*.{csv,ncu-rep,log}
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
LPS=$(python -c "import json,sys; print(json.loads([l for l in open('gpurun_out/${TAG}_plain.log') if l.startswith('{')][-1])['gpu_launches'])")
echo "launches per step: $LPS"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mstr:: \
    -s $((3 * LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_bf16_2cta -s $((31 * 3 + 18)) -c 1 \
    -o gpurun_out/${TAG}_conv $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
for K in conv3x3_flat tps_rectify bilstm_tc attn_context "rnn_gemm_kernel<0>" "rnn_gemm_kernel<1>" conv3x3_first maxpool2x2; do
  N=$(echo $K | tr -cd 'a-z0-9_')
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$K" -s 8 -c 1 \
      -o gpurun_out/${TAG}_k_$N $CMD > gpurun_out/${TAG}_ncu_$N.log 2>&1
  echo "ncu $K exit $?"
done
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --kernel-table gpurun_out/${TAG}_kernel_table.json > gpurun_out/${TAG}_bench.log 2>&1
tail -1 gpurun_out/${TAG}_bench.log > gpurun_out/${TAG}_bench_line.json
ls gpurun_out/ | grep ${TAG}_ | head -40

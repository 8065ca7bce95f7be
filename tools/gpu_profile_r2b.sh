#!/bin/bash
# ncu --set full captures of the kernels added / reworked in round 2 (each command has exited 0 without ncu before).
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$BENCH > gpurun_out/r02b_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 300 python tools/tune_seq.py 512 > gpurun_out/r02b_seq512.log 2>&1 || { echo "tune_seq failed"; exit 1; }
timeout 300 python tools/tf_one_forward.py 4096 > gpurun_out/r02b_tf.log 2>&1 || { echo "tf forward failed"; exit 1; }
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k regex:conv3x3_flat_pool -s 3 -c 1 -o gpurun_out/r02b_flat_pool $BENCH > gpurun_out/r02b_ncu1.log 2>&1; echo "flat_pool $?"
$NCU -k regex:attn_decode_persistent -s 3 -c 1 -o gpurun_out/r02b_decoder4096 $BENCH > gpurun_out/r02b_ncu2.log 2>&1; echo "decoder4096 $?"
$NCU -k regex:attn_decode_persistent -s 1 -c 1 -o gpurun_out/r02b_decoder512 python tools/tune_seq.py 512 > gpurun_out/r02b_ncu3.log 2>&1; echo "decoder512 $?"
$NCU -k regex:bilstm_cluster -s 1 -c 1 -o gpurun_out/r02b_bilstm_cluster512 python tools/tune_seq.py 512 > gpurun_out/r02b_ncu4.log 2>&1; echo "bilstm_cluster $?"
$NCU -k regex:mha_step -s 300 -c 1 -o gpurun_out/r02b_mha_step python tools/tf_one_forward.py 4096 > gpurun_out/r02b_ncu5.log 2>&1; echo "mha_step $?"
$NCU -k regex:mha_full -s 6 -c 1 -o gpurun_out/r02b_mha_full python tools/tf_one_forward.py 4096 > gpurun_out/r02b_ncu6.log 2>&1; echo "mha_full $?"
$NCU -k regex:layer_norm -s 488 -c 1 -o gpurun_out/r02b_layer_norm python tools/tf_one_forward.py 4096 > gpurun_out/r02b_ncu7.log 2>&1; echo "layer_norm $?"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mstr:: -s 156 -c 52 --csv --log-file gpurun_out/r02b_launches.csv $BENCH > gpurun_out/r02b_ncu8.log 2>&1; echo "launch list $?"
ls gpurun_out | grep r02b_ | head -30

#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02g.log
: > $L
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x -k "attn_decode_bf16 or bilstm or small_im2col or conv_gemm" >> $L 2>&1
echo "kernel tests exit $?" >> $L
timeout 300 python tools/tune_seq.py 256 512 1024 4096 >> $L 2>&1
echo "tune exit $?" >> $L
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_bf16_layers.py -m gpu -q -p no:cacheprovider -k "bf16" >> $L 2>&1
echo "model tests exit $?" >> $L
tail -40 $L

#!/bin/bash
# Run the GPU parity tests group by group, each in its own process with its own timeout, so that a
# trapped kernel (poisoned CUDA context) or a hang in one group does not hide the others' results.
# Usage (on the GPU box, from the repo root):  bash tools/gpu_checks.sh [bench]
set -u
mkdir -p gpurun_out
OUT=gpurun_out/checks.log
: > $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $OUT 2>&1
run() {  # name, timeout, command...
  local name=$1 tmo=$2; shift 2
  echo "=== $name" | tee -a $OUT
  timeout $tmo "$@" >> $OUT 2>&1
  local rc=$?
  echo "=== $name exit $rc" | tee -a $OUT
}
PT="python -m pytest -q -m gpu -p no:cacheprovider -x"
run tps        300 $PT tests/test_gpu_kernels.py -k "tps"
run first      300 $PT tests/test_gpu_kernels.py -k "conv3x3_first"
run conv_f32   300 $PT tests/test_gpu_kernels.py -k "conv_gemm_f32 or linear_f32"
run conv_tc    300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "bf16_tcgen05 or tf32x3"
run conv_tc2   300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "many_tiles or two_accumulators or fused_downsample"
run small      300 $PT tests/test_gpu_kernels.py -k "maxpool or mean or embed or errors"
run bilstm     300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "bilstm"
run attn       300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "attn and not bf16"
run attn_tc    300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "attn_decode_bf16"
run model_fp32 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py -k "fp32 or scene or tf32"
run model_bf16 600 python -m pytest -q -m gpu -p no:cacheprovider -s tests/test_gpu_model.py -k "bf16"
run smoke      300 python -c "import __graft_entry__ as g; g.smoke()"
run h2d        120 python tools/diag_h2d.py
run parity     600 python tests/parity_report.py 64
if [ "${1:-}" = "bench" ]; then
  run bench 900 python bench.py --steps 5 --warmup 3
  grep -h '^{' $OUT | tail -1 > gpurun_out/bench_line.json
fi
grep "^===" $OUT

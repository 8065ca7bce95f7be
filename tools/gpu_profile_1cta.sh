set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/p1_plain.log 2>&1 || { tail -5 gpurun_out/p1_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_bf16_kernel" -s 18 -c 6 -o gpurun_out/r01_conv_1cta $CMD > gpurun_out/p1_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/p1_ncu.log

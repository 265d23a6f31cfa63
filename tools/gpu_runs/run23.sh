cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_linear_head.py -q 2>&1 | tail -12 > gpurun_out/r3j_tests.log
timeout 300 python tools/quick_bench.py --Ks 3 5 10 12 16 --dtypes float32 bfloat16 float16 2>&1 | grep -E "grad" > gpurun_out/r3j_quick.log
for v in "" headlin1 headlin2; do
  if [ -n "$v" ]; then export BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so; fi
  echo "== head variant: ${v:-default}" >> gpurun_out/r3j_quick.log
  timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head" >> gpurun_out/r3j_quick.log
done

cd $GRAFT_REPO_ROOT
for v in "" t256 t256s3; do
  if [ -n "$v" ]; then export BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so; fi
  echo "== ${v:-default}" >> gpurun_out/r3s_quick.log
  timeout 300 python tools/quick_bench.py --Ks 1 2 3 4 5 --dtypes float32 bfloat16 2>&1 | grep -E "fwd\+grad" >> gpurun_out/r3s_quick.log
  timeout 300 python tools/quick_bench.py --T 64000 --Ks 1 2 5 --dtypes float32 2>&1 | grep -E "fwd\+grad" >> gpurun_out/r3s_quick.log
done

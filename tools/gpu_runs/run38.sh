cd $GRAFT_REPO_ROOT
for v in "" lin16 lin20 lin16m4; do
  if [ -n "$v" ]; then export BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so; fi
  echo "== ${v:-default}" >> gpurun_out/r4h_quick.log
  timeout 300 python tools/quick_bench.py --Ks 16 20 --dtypes bfloat16 float16 2>&1 | grep -E "fwd\+grad" >> gpurun_out/r4h_quick.log
done

cd $GRAFT_REPO_ROOT
for v in "" s3l1 s4l2 s4l2t256 s3l2t256; do
  if [ -n "$v" ]; then export BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so; fi
  echo "== ${v:-default}" >> gpurun_out/r4d_quick.log
  timeout 300 python tools/quick_bench.py --Ks 1 2 3 5 --dtypes float32 2>&1 | grep -E "fwd\+grad" >> gpurun_out/r4d_quick.log
done

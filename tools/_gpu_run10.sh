cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_linear_head.py -q 2>&1 | tail -3 > gpurun_out/r2o_linear_tests.log
echo "== minb 6 default" >> gpurun_out/r2o_linear_perf.log; timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head" >> gpurun_out/r2o_linear_perf.log
for mb in 5 7 8; do echo "== minb $mb" >> gpurun_out/r2o_linear_perf.log; BLVM_B200_LINEAR_CTAS_PER_SM=$mb BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_lin$mb.so timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head" >> gpurun_out/r2o_linear_perf.log; done

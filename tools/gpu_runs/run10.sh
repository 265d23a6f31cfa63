cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_linear_head.py -q 2>&1 | tail -3 > gpurun_out/r2t_linear_tests.log
echo "== default (DIN=30 specialised)" >> gpurun_out/r2t_linear_perf.log; timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head" >> gpurun_out/r2t_linear_perf.log
echo "== backoff" >> gpurun_out/r2t_linear_perf.log; BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_linbo.so timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head" >> gpurun_out/r2t_linear_perf.log

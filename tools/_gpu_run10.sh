cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_linear_head.py -q -x 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2j_linear_tests.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2j_all_tests.log
for n in 3 5 7 8; do echo "== ctas/sm $n" >> gpurun_out/r2j_linear_perf.log; BLVM_B200_LINEAR_CTAS_PER_SM=$n timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "B=256|fused head" >> gpurun_out/r2j_linear_perf.log; done

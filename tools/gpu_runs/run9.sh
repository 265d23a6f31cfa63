cd $GRAFT_REPO_ROOT
for n in 0 4 5; do echo "== ctas/sm $n" >> gpurun_out/r2i_linear3.log; BLVM_B200_DEBUG=1 BLVM_B200_LINEAR_CTAS_PER_SM=$n timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "B=|fused head|occupancy|unfused|rc=|Error|error" >> gpurun_out/r2i_linear3.log; done
timeout 180 python tools/test_linear_dmol.py --fp16 2>&1 | grep -E "B=|fused head" >> gpurun_out/r2i_linear3.log

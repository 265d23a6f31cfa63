cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_reference_models.py tests/test_gpu_linear_head.py -q -s 2>&1 | grep -E "^\[|passed|failed|Error|^E  |FAILED" | cut -c1-400 > gpurun_out/r2p_models.log
timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head|unfused" > gpurun_out/r2p_linear_perf.log

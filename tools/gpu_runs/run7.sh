cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_reference_models.py -q -s -k "fp32" 2>&1 | grep -E "^\[|passed|failed|^E  " | cut -c1-330 > gpurun_out/r2g_models.log

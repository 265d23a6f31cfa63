cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_reference_models.py -q -s -k "cwvae or sync or nansum" 2>&1 | grep -v "Warning\|warnings.warn\|^tests/test_gpu_ref.*warnings$" | cut -c1-400 | tail -150 > gpurun_out/r2c_pytest.log

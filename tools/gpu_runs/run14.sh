cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_linear_head.py -q -x 2>&1 | tail -8 > gpurun_out/r2z_tests.log
timeout 300 python tools/quick_bench.py --Ks 1 5 8 10 16 30 --dtypes float32 bfloat16 2>&1 | grep -E "grad" > gpurun_out/r2z_quick.log
timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head" >> gpurun_out/r2z_quick.log

cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_linear_head.py -q 2>&1 | tail -12 > gpurun_out/r3i_tests.log
timeout 300 python tools/quick_bench.py --Ks 2 5 8 10 16 30 --dtypes float32 bfloat16 2>&1 | grep -E "grad" > gpurun_out/r3i_quick.log
timeout 180 python tools/test_linear_dmol.py 2>&1 | grep -E "fused head" >> gpurun_out/r3i_quick.log

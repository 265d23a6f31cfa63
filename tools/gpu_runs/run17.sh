cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x 2>&1 | tail -5 > gpurun_out/r3c_tests.log
timeout 300 python tools/quick_bench.py --Ks 1 2 3 4 5 8 10 16 30 --dtypes float32 bfloat16 2>&1 | grep -E "grad" > gpurun_out/r3c_quick.log

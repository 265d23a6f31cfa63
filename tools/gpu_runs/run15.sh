cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x 2>&1 | tail -5 > gpurun_out/r3a_tests.log
timeout 300 python tools/quick_bench.py --Ks 1 2 3 4 5 --dtypes float32 bfloat16 2>&1 | grep -E "grad|fwd only" > gpurun_out/r3a_quick.log
timeout 300 python tools/quick_bench.py --T 64000 --Ks 1 2 5 --dtypes float32 2>&1 | grep -E "grad|fwd only" >> gpurun_out/r3a_quick.log

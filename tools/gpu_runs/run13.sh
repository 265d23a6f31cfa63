cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_linear_head.py -q -x 2>&1 | tail -8 > gpurun_out/r2y_tests.log
timeout 300 python tools/quick_bench.py --Ks 8 12 16 20 --dtypes float32 bfloat16 2>&1 | grep -E "grad|fwd only" > gpurun_out/r2y_quick.log
timeout 300 python tools/head_step_profile.py 2>&1 | grep -E "blvm::|fuse_linear|Self CUDA time" > gpurun_out/r2y_headprof.log

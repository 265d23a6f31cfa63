cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3 > gpurun_out/r3d_tests.log
timeout 300 python tools/quick_bench.py --Ks 10 16 20 30 --dtypes float32 bfloat16 2>&1 | grep -E "grad" > gpurun_out/r3d_quick.log

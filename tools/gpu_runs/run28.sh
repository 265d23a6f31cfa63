cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -4 > gpurun_out/r3r_tests.log
timeout 300 python tools/quick_bench.py --Ks 2 3 4 5 --dtypes float32 2>&1 | grep -E "grad|fwd only" > gpurun_out/r3r_quick.log

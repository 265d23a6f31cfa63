cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r3t_tests.log
timeout 300 python tools/quick_bench.py --Ks 2 3 4 5 --dtypes float32 bfloat16 2>&1 | grep -E "fwd\+grad|fwd only" > gpurun_out/r3t_quick.log

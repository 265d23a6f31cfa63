cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q 2>&1 | tail -6 > gpurun_out/r3m_tests.log
timeout 300 python tools/quick_bench.py --Ks 10 12 --dtypes bfloat16 float16 2>&1 | grep -E "grad" > gpurun_out/r3m_quick.log
timeout 300 python tools/quick_bench.py --T 64000 --Ks 10 --dtypes bfloat16 float16 2>&1 | grep -E "grad" >> gpurun_out/r3m_quick.log

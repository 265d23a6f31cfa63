cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3v_tests.log
timeout 600 python bench.py --dtype f16 --no-cpu-baseline --no-reference-cuda --no-e2e --no-sweep > gpurun_out/r3v_bench_n1_f16.json 2> gpurun_out/r3v_bench_n1_f16.err

cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_reference_models.py -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert " | cut -c1-700 > gpurun_out/r2d_models.log
timeout 300 python tools/quick_bench.py --Ks 8 10 16 30 --dtypes float32 bfloat16 2>&1 | grep dmol > gpurun_out/r2d_quick.log
for v in smk12 unc g3; do BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so timeout 300 python tools/quick_bench.py --Ks 8 10 12 --dtypes float32 bfloat16 2>&1 | grep dmol > gpurun_out/r2d_quick_$v.log; done
timeout 900 python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
timeout 300 python bench.py --dtype bf16 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2d_bench_n1_bf16.json 2> gpurun_out/r2d_bench_n1_bf16.err
timeout 300 python bench.py --B 32 --no-cpu-baseline --no-reference-cuda --no-sweep --no-e2e > gpurun_out/r2d_bench_b32.json 2> gpurun_out/r2d_bench_b32.err
tail -3 gpurun_out/r2d_bench_n1.err

cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert |^FAILED|^E  " | cut -c1-600 > gpurun_out/r2e_pytest.log
BLVM_B200_FUSED_MAX_TILES=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q 2>&1 | tail -3 > gpurun_out/r2e_pytest_unfused.log
for w in config2 config3 config4; do
  timeout 200 python bench.py --workload $w --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda > gpurun_out/r2e_bench_$w.json 2> gpurun_out/r2e_bench_$w.err
  BLVM_B200_FUSED_MAX_TILES=0 timeout 200 python bench.py --workload $w --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda > gpurun_out/r2e_bench_${w}_unfused.json 2>> gpurun_out/r2e_bench_$w.err
  BLVM_B200_FUSED_MAX_TILES=100000 timeout 200 python bench.py --workload $w --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda > gpurun_out/r2e_bench_${w}_forcefused.json 2>> gpurun_out/r2e_bench_$w.err
done
for b in 32 64 128; do
  timeout 200 python bench.py --B $b --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep > gpurun_out/r2e_bench_b$b.json 2> gpurun_out/r2e_bench_b$b.err
  BLVM_B200_FUSED_MAX_TILES=100000 timeout 200 python bench.py --B $b --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep > gpurun_out/r2e_bench_b${b}_forcefused.json 2>> gpurun_out/r2e_bench_b$b.err
  BLVM_B200_FUSED_MAX_TILES=0 timeout 200 python bench.py --B $b --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep > gpurun_out/r2e_bench_b${b}_unfused.json 2>> gpurun_out/r2e_bench_b$b.err
done
timeout 200 python bench.py --workload config2 --mode eager --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda > gpurun_out/r2e_bench_config2_eager.json 2>> gpurun_out/r2e_bench_config2.err
timeout 200 python bench.py --workload config4 --mode eager --steps 300 --no-e2e --no-cpu-baseline --no-reference-cuda > gpurun_out/r2e_bench_config4_eager.json 2>> gpurun_out/r2e_bench_config4.err
for v in smk12 unc g3; do BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so timeout 300 python tools/quick_bench.py --Ks 8 10 12 --dtypes float32 bfloat16 > gpurun_out/r2e_quick_$v.log 2>&1; done
timeout 200 python tools/quick_bench.py --Ks 1 2 3 4 5 --dtypes float32 bfloat16 > gpurun_out/r2e_quick_smallk.log 2>&1

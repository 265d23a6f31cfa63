cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_reference_models.py -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert |^FAILED|^E  " | cut -c1-500 > gpurun_out/r2f_models.log
for v in smk12 unc g3; do BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so timeout 300 python tools/quick_bench.py --Ks 8 10 12 --dtypes float32 bfloat16 2>&1 | grep -i "dmol\|error" > gpurun_out/r2f_quick_$v.log; done
# launch list of the bench step (after the same command exited 0 without ncu)
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r2f_bench_short.json 2> gpurun_out/r2f_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r2f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmol_tile_kernel -c 1 -o gpurun_out/r2f_dmol_k10_f32 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r2f_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_dmol_k10_f32.ncu-rep > gpurun_out/r2f_ncu_dmol_k10_f32.summary.json 2>&1; rm -f gpurun_out/r2f_dmol_k10_f32.ncu-rep
python bench.py --dtype bf16 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r2f_bench_short_bf16.json 2>> gpurun_out/r2f_bench_short.err && \
ncu --set full --clock-control none --import-source on -k regex:dmol_tile_kernel -c 1 -o gpurun_out/r2f_dmol_k10_bf16 python bench.py --dtype bf16 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r2f_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_dmol_k10_bf16.ncu-rep > gpurun_out/r2f_ncu_dmol_k10_bf16.summary.json 2>&1; rm -f gpurun_out/r2f_dmol_k10_bf16.ncu-rep
ncu --set full --clock-control none -k regex:"kl_multi_kernel|elbo_finalize" -c 2 -o gpurun_out/r2f_kl_fin python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r2f_ncu3.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_kl_fin.ncu-rep > gpurun_out/r2f_ncu_kl_finalize.summary.json 2>&1; rm -f gpurun_out/r2f_kl_fin.ncu-rep
ls gpurun_out | grep r2f

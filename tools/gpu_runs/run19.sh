cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r3e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3e_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r3e_bench_n1.json 2> gpurun_out/r3e_bench_n1.err
timeout 600 python bench.py --dtype bf16 --no-cpu-baseline --no-reference-cuda > gpurun_out/r3e_bench_n1_bf16.json 2> gpurun_out/r3e_bench_n1_bf16.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r3e_bench_reference.json 2> gpurun_out/r3e_bench_reference.err
# launch list of the bench step (after the same command exited 0 without ncu)
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r3e_bench_short.json 2> gpurun_out/r3e_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3e_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r3e_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmol_tile_kernel -c 1 -o gpurun_out/r3e_dmol_k10_f32 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r3e_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/r3e_dmol_k10_f32.ncu-rep > gpurun_out/r3e_ncu_dmol_k10_f32.summary.json 2>&1; rm -f gpurun_out/r3e_dmol_k10_f32.ncu-rep
python bench.py --dtype bf16 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r3e_bench_short_bf16.json 2>> gpurun_out/r3e_bench_short.err && \
ncu --set full --clock-control none --import-source on -k regex:dmol_tile_kernel -c 1 -o gpurun_out/r3e_dmol_k10_bf16 python bench.py --dtype bf16 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r3e_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/r3e_dmol_k10_bf16.ncu-rep > gpurun_out/r3e_ncu_dmol_k10_bf16.summary.json 2>&1
ncu -i gpurun_out/r3e_dmol_k10_bf16.ncu-rep --page source --csv > gpurun_out/r3e_bf16_k10.source.csv 2>/dev/null; rm -f gpurun_out/r3e_dmol_k10_bf16.ncu-rep
ncu --set full --clock-control none -k regex:"kl_multi_kernel|elbo_finalize" -c 2 -o gpurun_out/r3e_kl_fin python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r3e_ncu3.log 2>&1
python tools/ncu_summary.py gpurun_out/r3e_kl_fin.ncu-rep > gpurun_out/r3e_ncu_kl_finalize.summary.json 2>&1; rm -f gpurun_out/r3e_kl_fin.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:linear_dmol_kernel -c 1 -o gpurun_out/r3e_head python tools/test_linear_dmol.py > gpurun_out/r3e_ncu4.log 2>&1
python tools/ncu_summary.py gpurun_out/r3e_head.ncu-rep > gpurun_out/r3e_ncu_linear_dmol_head_bf16.summary.json 2>&1
ncu -i gpurun_out/r3e_head.ncu-rep --page source --csv > gpurun_out/r3e_head.source.csv 2>/dev/null; rm -f gpurun_out/r3e_head.ncu-rep
ncu --set full --clock-control none -k regex:dmol_tile_kernel -c 1 -o gpurun_out/r3e_k16 python tools/quick_bench.py --Ks 16 --dtypes float32 > gpurun_out/r3e_ncu5.log 2>&1
python tools/ncu_summary.py gpurun_out/r3e_k16.ncu-rep > gpurun_out/r3e_ncu_dmol_k16_f32.summary.json 2>&1; rm -f gpurun_out/r3e_k16.ncu-rep
ls gpurun_out | grep r3e

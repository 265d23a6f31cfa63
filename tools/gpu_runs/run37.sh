cd $GRAFT_REPO_ROOT
python bench.py --dtype f16 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r4e_bench_short_f16.json 2> gpurun_out/r4e_err.log && \
ncu --set full --clock-control none -k regex:dmol_tile_kernel -c 1 -o gpurun_out/r4e_f16 python bench.py --dtype f16 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-cuda --no-sweep --min-seconds 0 > gpurun_out/r4e_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/r4e_f16.ncu-rep > gpurun_out/r4e_ncu_dmol_k10_f16.summary.json 2>&1; rm -f gpurun_out/r4e_f16.ncu-rep
ncu --set full --clock-control none -k regex:dmol_stream_kernel -s 8 -c 1 -o gpurun_out/r4e_k5 python tools/quick_bench.py --Ks 5 --dtypes float32 > gpurun_out/r4e_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/r4e_k5.ncu-rep > gpurun_out/r4e_ncu_stream_k5_f32.summary.json 2>&1; rm -f gpurun_out/r4e_k5.ncu-rep

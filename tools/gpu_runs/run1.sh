set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v Warning | tail -60 > gpurun_out/r2a_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python tools/quick_bench.py --Ks 2 5 10 16 30 --dtypes float32 bfloat16 float16 > gpurun_out/r2a_quick.log 2>&1
for w in config2 config4 config5; do timeout 120 python tools/host_overhead.py $w 2>&1 | head -40 > gpurun_out/r2a_host_$w.log; done
for w in config2 config3 config4; do timeout 200 python bench.py --workload $w --steps 300 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dmol_tile_kernel -c 2 -o gpurun_out/r2a_bf16_k10 python tools/quick_bench.py --Ks 10 --dtypes bfloat16 > gpurun_out/r2a_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2a_bf16_k10.ncu-rep > gpurun_out/r2a_ncu_bf16_k10.summary.json 2>&1
ncu -i gpurun_out/r2a_bf16_k10.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/r2a_bf16_k10.source.csv 2>/dev/null
ls -la gpurun_out

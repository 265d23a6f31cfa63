set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -v Warning | tail -80 > gpurun_out/r2b_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python tools/quick_bench.py --Ks 6 8 10 12 16 20 30 --dtypes float32 bfloat16 float16 > gpurun_out/r2b_quick.log 2>&1
for v in gf2 gh4 gh1; do BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so timeout 300 python tools/quick_bench.py --Ks 8 10 16 30 --dtypes float32 bfloat16 > gpurun_out/r2b_quick_$v.log 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dmol_tile_kernel -c 1 -o gpurun_out/r2b_bf16_k10 python tools/quick_bench.py --Ks 10 --dtypes bfloat16 > gpurun_out/r2b_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2b_bf16_k10.ncu-rep > gpurun_out/r2b_ncu_bf16_k10.summary.json 2>&1
ncu -i gpurun_out/r2b_bf16_k10.ncu-rep --page source --csv > gpurun_out/r2b_bf16_k10.source.csv 2>/dev/null
rm -f gpurun_out/r2b_bf16_k10.ncu-rep
ls -la gpurun_out | tail -12

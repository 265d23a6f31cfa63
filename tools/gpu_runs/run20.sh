cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:dmol_stream_kernel -s 8 -c 1 -o gpurun_out/r3g_stream_k2 python tools/quick_bench.py --Ks 2 --dtypes float32 > gpurun_out/r3g_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/r3g_stream_k2.ncu-rep > gpurun_out/r3g_stream_k2.summary.json 2>&1
ncu -i gpurun_out/r3g_stream_k2.ncu-rep --page source --csv > gpurun_out/r3g_stream_k2.source.csv 2>/dev/null
rm -f gpurun_out/r3g_stream_k2.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:linear_dmol_kernel -c 8 -o gpurun_out/r3g_head python tools/test_linear_dmol.py > gpurun_out/r3g_ncu4.log 2>&1
python tools/ncu_summary.py gpurun_out/r3g_head.ncu-rep > gpurun_out/r3g_head_all.summary.json 2>&1
rm -f gpurun_out/r3g_head.ncu-rep

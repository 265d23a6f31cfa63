cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:dmol_stream_kernel -s 8 -c 1 -o gpurun_out/r3b_stream_k5 python tools/quick_bench.py --Ks 5 --dtypes float32 > gpurun_out/r3b_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/r3b_stream_k5.ncu-rep > gpurun_out/r3b_stream_k5.summary.json 2>&1
ncu -i gpurun_out/r3b_stream_k5.ncu-rep --page source --csv > gpurun_out/r3b_stream_k5.source.csv 2>/dev/null
ncu -i gpurun_out/r3b_stream_k5.ncu-rep --page details --csv > gpurun_out/r3b_stream_k5.details.csv 2>/dev/null
rm -f gpurun_out/r3b_stream_k5.ncu-rep

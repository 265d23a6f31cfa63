cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:linear_dmol_kernel -s 5 -c 1 -o gpurun_out/r3h_head python tools/test_linear_dmol.py > gpurun_out/r3h_ncu.log 2>&1
ncu -i gpurun_out/r3h_head.ncu-rep --page source --csv > gpurun_out/r3h_head.source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r3h_head.ncu-rep > gpurun_out/r3h_head.summary.json 2>&1
rm -f gpurun_out/r3h_head.ncu-rep

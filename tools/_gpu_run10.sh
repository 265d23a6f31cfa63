cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:linear_dmol_kernel -s 8 -c 1 -o gpurun_out/r2q_linear python tools/test_linear_dmol.py > gpurun_out/r2q_ncu.log 2>&1
ncu -i gpurun_out/r2q_linear.ncu-rep --page source --csv > gpurun_out/r2q_linear.source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2q_linear.ncu-rep > gpurun_out/r2q_linear.summary.json 2>&1
rm -f gpurun_out/r2q_linear.ncu-rep

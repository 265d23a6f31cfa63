cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r4c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4c_smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r4c_bench_reference.json 2> gpurun_out/r4c_bench_reference.err
timeout 900 python bench.py > gpurun_out/r4c_bench_n1.json 2> gpurun_out/r4c_bench_n1.err

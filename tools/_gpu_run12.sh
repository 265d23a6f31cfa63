cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2r_pytest.log
timeout 900 python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err
timeout 300 python bench.py --dtype bf16 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2r_bench_n1_bf16.json 2> gpurun_out/r2r_bench_n1_bf16.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1
tail -2 gpurun_out/r2r_bench_n1.err

cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/r3u_bench_n1.json 2> gpurun_out/r3u_bench_n1.err
timeout 600 python bench.py --dtype bf16 --no-cpu-baseline --no-reference-cuda > gpurun_out/r3u_bench_n1_bf16.json 2> gpurun_out/r3u_bench_n1_bf16.err
timeout 600 python bench.py --dtype f16 --no-cpu-baseline --no-reference-cuda --no-e2e > gpurun_out/r3u_bench_n1_f16.json 2> gpurun_out/r3u_bench_n1_f16.err

cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 > gpurun_out/r3w_multi_tests.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 2 --steps 200 --warmup 10 --dtype f16 --no-e2e --no-cpu-baseline > gpurun_out/r3w_bench_n2_f16.json 2> gpurun_out/r3w_bench_n2_f16.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3w_bench_n2_ref.json 2> gpurun_out/r3w_bench_n2_ref.err

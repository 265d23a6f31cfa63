cd $GRAFT_REPO_ROOT
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r3n_bench_n$N.json 2> gpurun_out/r3n_bench_n$N.err
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 > gpurun_out/r3n_multi_tests.log; fi

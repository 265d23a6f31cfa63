cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r2v_bench_n8.json 2> gpurun_out/r2v_bench_n8.err

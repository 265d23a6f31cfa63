cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s 2>&1 | tail -15 > gpurun_out/r2h_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --scaling strong --no-e2e > gpurun_out/r2h_bench_n2_strong.json 2>> gpurun_out/r2h_bench_n2.err
tail -5 gpurun_out/r2h_bench_n2.err

cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2s_topo8.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 200 --warmup 10 --no-e2e > gpurun_out/r2s_bench_n4.json 2> gpurun_out/r2s_bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 200 --warmup 10 --B 512 --scaling strong --no-e2e > gpurun_out/r2s_bench_n8_strong512.json 2> gpurun_out/r2s_bench_n8_strong512.err
tail -3 gpurun_out/r2s_bench_n8.err

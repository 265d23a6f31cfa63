cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5 > gpurun_out/r2u_multi_tests.log
for w in config3 config4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 10 --workload $w --no-sweep --no-reference-cuda > gpurun_out/r2u_bench_n2_$w.json 2> gpurun_out/r2u_bench_n2_$w.err
done

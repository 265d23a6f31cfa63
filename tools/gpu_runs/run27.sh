cd $GRAFT_REPO_ROOT
for v in "" exNOFENCE exNOCONSUME; do
  if [ -n "$v" ]; then export BLVM_B200_LIB=$PWD/benchmarking-lvms_b200/lib/variants/libblvm_b200_$v.so; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 300 --warmup 10 --no-strong --no-e2e --no-cpu-baseline > gpurun_out/r3q_n2_${v:-default}.json 2> gpurun_out/r3q_n2_${v:-default}.err
done
unset BLVM_B200_LIB
timeout 200 python bench.py --steps 300 --warmup 10 --no-e2e --no-cpu-baseline --no-sweep --no-reference-cuda > gpurun_out/r3q_n1.json 2> gpurun_out/r3q_n1.err

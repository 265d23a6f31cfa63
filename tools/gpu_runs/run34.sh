cd $GRAFT_REPO_ROOT
timeout 1500 python tools/model_step_bench.py > gpurun_out/r3y_model_steps.jsonl 2> gpurun_out/r3y_model_steps.err

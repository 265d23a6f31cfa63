"""Development aid: host time spent inside blvm_b200 during one patched reference-model training step (cProfile)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import model_step_bench as msb  # noqa: E402  (loads the staged reference)

B = msb.B
name = sys.argv[1] if len(sys.argv) > 1 else "config3_srnn"
for variant in ("reference", "patched"):
    if variant == "patched":
        B.patch_blvm()
    model, Bn, T, kw = msb.build(name)
    model = model.cuda().train()
    g = torch.Generator().manual_seed(1)
    x = (torch.randint(0, msb.NB, (Bn, T), generator=g).float() / (msb.NB - 1) * 2 - 1).cuda()
    x_sl = torch.full((Bn,), T, dtype=torch.int64)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    scaler = torch.amp.GradScaler("cuda")
    for _ in range(4):
        msb.step(model, opt, scaler, x, x_sl, kw)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        msb.step(model, opt, scaler, x, x_sl, kw)
    torch.cuda.synchronize()
    pr.disable()
    print("=====", name, variant)
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(r"benchmarking-lvms_b200|model_step_bench|autograd/__init__|_tensor.py", 12)
    st.sort_stats("tottime").print_stats(22)
    if variant == "patched":
        B.unpatch_blvm()

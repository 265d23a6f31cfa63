"""Development aid: which kernels run in one fused-head training step through the public API (torch.profiler table)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blvm_b200  # noqa: E402

dev = torch.device("cuda", 0)
B, T, K = 256, 16000, 10
NB = 65536
Din = 3 * K
y = (torch.randint(0, NB, (B, T), device=dev).float() / (NB - 1) * 2 - 1)
x = torch.randn(B, T, Din, device=dev).to(torch.bfloat16).requires_grad_(True)
x_sl = torch.full((B,), T, dtype=torch.int64)
x_dev = x_sl.to(dev)
for fuse in (True, False):
    lik = blvm_b200.DiscretizedLogisticMixtureDense(Din, 1, K, NB, fuse_linear=fuse).to(dev)

    def step():
        x.grad = None
        lik.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            params = lik(x)
        r = blvm_b200.fused_elbo(y, params, x_sl, (), num_bins=NB, denom=float(B * T), x_sl_device=x_dev)
        r.loss.backward()
        return r

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    print("fuse_linear =", fuse)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=90))

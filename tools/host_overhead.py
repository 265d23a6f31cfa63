"""Development aid: CPU enqueue time per step of the eager fused ELBO path (vs GPU time), with a cProfile of the host side."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blvm_b200  # noqa: E402

dev = "cuda"
SHAPES = {"config5": (256, 16000, [(64, 64)]), "config2": (32, 16000, []), "config3": (64, 32000, [(64, 64)]),
          "config4": (32, 65536, [(64, 128), (512, 64), (4096, 32)]), "tiny": (4, 2048, [(64, 16)])}
B, T, LEVELS = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "config5"]
K, nb = 10, 65536
y = torch.rand(B, T, device=dev) * 2 - 1
raw = torch.randn(B, T, 3 * K, device=dev, requires_grad=True)
kls = [[torch.randn(B, T // S, Z, device=dev, requires_grad=True), (torch.rand(B, T // S, Z, device=dev) + 0.1).requires_grad_(True),
        torch.randn(B, T // S, Z, device=dev, requires_grad=True), (torch.rand(B, T // S, Z, device=dev) + 0.1).requires_grad_(True)]
       for S, Z in LEVELS]
x_sl = torch.full((B,), T)
x_dev = x_sl.to(dev)
lens = [blvm_b200.level_lengths(x_dev, S) for S, _ in LEVELS]
params = blvm_b200.DMoLParams(raw, K, 1, -7.0)
print(f"shape {sys.argv[1] if len(sys.argv) > 1 else 'config5'}: B={B} T={T} levels={LEVELS}")


def step(host_lengths=False):
    raw.grad = None
    for kl in kls:
        for t in kl:
            t.grad = None
    if host_lengths:
        out = blvm_b200.fused_elbo(y, params, x_sl, [blvm_b200.KLLevel(*kl, stride=S) for kl, (S, _) in zip(kls, LEVELS)], 0.5, 0.0625, num_bins=nb)
    else:
        out = blvm_b200.fused_elbo(y, params, x_sl, [blvm_b200.KLLevel(*kl, lens=ln) for kl, ln in zip(kls, lens)], 0.5, 0.0625, num_bins=nb,
                                   denom=float(B * T), x_sl_device=x_dev)
    out.loss.backward()


for hl in (False, True):
    for _ in range(20):
        step(hl)
    torch.cuda.synchronize()
    # tiny problem so the GPU never back-pressures: pure host cost
    N = 300
    t0 = time.perf_counter()
    for _ in range(N):
        step(hl)
    t_enq = (time.perf_counter() - t0) / N
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / N
    print(f"host_lengths={hl}: enqueue {t_enq * 1e6:.1f} us/step, wall incl. GPU {t_all * 1e6:.1f} us/step")

pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step(False)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

"""Device-only timing of the individual kernels (development aid; bench.py is the contract benchmark)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blvm_b200  # noqa: E402
from blvm_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    # back-to-back launches between ONE event pair (what bench.py reports): no per-launch event / idle-gap overhead
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--T", type=int, default=16000)
    ap.add_argument("--Ks", type=int, nargs="+", default=[10])
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--dtypes", nargs="+", default=["float32"])
    args = ap.parse_args()
    dev = "cuda"
    B, T = args.B, args.T
    nb = 65536
    a = torch.empty(1 << 28, device=dev)  # 1 GiB
    b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a))
    print(f"copy 1GiB fp32: {2 * a.numel() * 4 / best / 1e6:.0f} GB/s best, {2 * a.numel() * 4 / med / 1e6:.0f} median")
    del a, b
    for K, dtn in [(K, d) for K in args.Ks for d in args.dtypes]:
        dt = getattr(torch, dtn)
        esz = torch.empty(0, dtype=dt).element_size()
        y = torch.randint(0, nb, (B, T), device=dev).float() / (nb - 1) * 2 - 1
        raw = torch.randn(B, T, 3 * K, device=dev)
        raw[..., K:2 * K] = y.unsqueeze(-1) + 0.1 * torch.randn(B, T, K, device=dev)
        raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
        raw = raw.to(dt)
        x_sl = torch.full((B,), T, dtype=torch.int64)
        if args.ragged:
            x_sl = (T * (0.5 + 0.5 * torch.rand(B))).long()
        x_dev = x_sl.to(dev)
        lp = torch.empty(B, T, device=dev)
        graw = torch.empty_like(raw)
        part = torch.empty(B * int(blvm_b200._lib.lib.blvm_dmol_chunks(T, K, 1)), dtype=torch.float64, device=dev)
        N = B * T
        g = -1.0 / float(x_sl.sum())
        med, best = timeit(lambda: ops._dmol_call(y, raw, x_dev, None, g, B, T, K, 1, nb, -7.0, 1, lp, graw, part))
        byt = N * (8 + 6 * K * esz)
        print(f"K={K:2d} {dtn:8s} dmol fwd+grad: {med * 1e3:8.1f} us loop ({best * 1e3:.1f} best)  {N / med / 1e6:7.2f} Gsamples/s  "
              f"{byt / med / 1e6:7.0f} GB/s algorithmic")
        med, best = timeit(lambda: ops._dmol_call(y, raw, x_dev, None, 0.0, B, T, K, 1, nb, -7.0, 1, lp, None, part))
        byt = N * (8 + 3 * K * esz)
        print(f"K={K:2d} {dtn:8s} dmol fwd only: {med * 1e3:8.1f} us loop ({best * 1e3:.1f} best)  {N / med / 1e6:7.2f} Gsamples/s  "
              f"{byt / med / 1e6:7.0f} GB/s algorithmic")
        med, best = timeit(lambda: ops.dmol_sample_mode(raw, K, 1, -7.0))
        print(f"K={K:2d} {dtn:8s} sample+mode  : {med * 1e3:8.1f} us loop ({best * 1e3:.1f} best)  {N / med / 1e6:7.2f} Gsamples/s  "
              f"{N * (3 * K * esz + 12) / med / 1e6:7.0f} GB/s if every parameter byte were read")
        del raw, graw
    # single discretized logistic (DiscretizedLogisticDense): packed (B, T, 2)
    y = torch.randint(0, nb, (B, T), device=dev).float() / (nb - 1) * 2 - 1
    raw2 = torch.randn(B, T, 2, device=dev)
    raw2[..., 0] = y + 0.1 * raw2[..., 0]
    raw2[..., 1] = raw2[..., 1] * 2 - 4
    x_dev = torch.full((B,), T, dtype=torch.int64, device=dev)
    lp = torch.empty(B, T, device=dev)
    g2 = torch.empty_like(raw2)
    part = torch.empty(B * int(blvm_b200._lib.lib.blvm_dl_chunks(T)), dtype=torch.float64, device=dev)
    med, best = timeit(lambda: ops._dl_call(y, raw2, x_dev, None, -1e-6, B, T, nb, -7.0, 1, lp, g2, part))
    print(f"DL fwd+grad: {med * 1e3:8.1f} us loop; {B * T * 24 / med / 1e6:7.0f} GB/s algorithmic")
    med, best = timeit(lambda: ops._dl_call(y, raw2, x_dev, None, 0.0, B, T, nb, -7.0, 1, lp, None, part))
    print(f"DL fwd only: {med * 1e3:8.1f} us loop; {B * T * 16 / med / 1e6:7.0f} GB/s algorithmic")
    # KL fused
    S, Z = 64, 64
    Tz = T // S
    ins = [torch.randn(B, Tz, Z, device=dev), torch.rand(B, Tz, Z, device=dev) + 0.1, torch.randn(B, Tz, Z, device=dev),
           torch.rand(B, Tz, Z, device=dev) + 0.1]
    for t in ins:
        t.requires_grad_(True)
    y = torch.zeros(B, T, device=dev)

    def kl_step():
        out = blvm_b200.fused_elbo(None, None, torch.full((B,), T), [blvm_b200.KLLevel(*ins, stride=S)], 0.5, 0.0625, num_bins=2)
        return out

    med, best = timeit(kl_step)
    L = B * Tz * Z
    print(f"KL fused (+finalize, python): {med * 1e3:.1f} us loop; {32 * L / med / 1e6:.0f} GB/s algorithmic (L={L})")


if __name__ == "__main__":
    main()

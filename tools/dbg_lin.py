import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, blvm_b200
from blvm_b200 import ops
lib = blvm_b200._lib.lib
import sys as _s
K, Bn, T, nb = int(_s.argv[1]), 3, 4096, 65536
gen = torch.Generator().manual_seed(K * 1000 + T)
y = (torch.randint(0, nb, (Bn, T), generator=gen).float() / (nb - 1) * 2 - 1).cuda()
raw = torch.randn(Bn, T, 3 * K, generator=gen)
raw[..., K:2 * K] = y.cpu().unsqueeze(-1) + 0.1 * raw[..., K:2 * K]
raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
raw = raw.to(torch.bfloat16).cuda()
x_dev = torch.full((Bn,), T, dtype=torch.int64).cuda()
chunks = int(lib.blvm_dmol_chunks(T, K, 1))
for mode in (0, 1):
    lib.blvm_set_stream_mode(mode)
    lp = torch.zeros(Bn, T, device="cuda"); graw = torch.zeros_like(raw); part = torch.zeros(Bn * chunks, dtype=torch.float64, device="cuda")
    ops._dmol_call(y, raw, x_dev, None, -0.37, Bn, T, K, 1, nb, -7.0, 1, lp, graw, part)
    lp2 = torch.zeros(Bn, T, device="cuda"); part2 = torch.zeros_like(part)
    ops._dmol_call(y, raw, x_dev, None, 0.0, Bn, T, K, 1, nb, -7.0, 1, lp2, None, part2)
    torch.cuda.synchronize()
    d = (lp != lp2)
    print("mode", mode, "differing", int(d.sum()), "of", lp.numel())
    idx = d.nonzero()[:0]
    for b, t in idx.tolist():
        print("  ", b, t, float(lp[b, t]), float(lp2[b, t]), float(lp[b, t] - lp2[b, t]), "y", float(y[b, t]), "raw", raw[b, t].float().tolist())
    rf = raw.float()
    lpf = torch.zeros(Bn, T, device="cuda"); gf = torch.zeros_like(rf); pf = torch.zeros_like(part)
    ops._dmol_call(y, rf, x_dev, None, -0.37, Bn, T, K, 1, nb, -7.0, 1, lpf, gf, pf)
    torch.cuda.synchronize()
    print("   grad-kernel == fp32(log-domain) at differing:", int((lp[d] == lpf[d]).sum()), " fwd-kernel == fp32 at differing:", int((lp2[d] == lpf[d]).sum()),
          " overall grad==fp32:", int((lp == lpf).sum()), " fwd==fp32:", int((lp2 == lpf).sum()))

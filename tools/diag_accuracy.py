"""Development aid: where is the CUDA path furthest from the fp64 oracle? (per parameter group)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blvm_b200  # noqa: E402
from oracle import blvm_oracle as O  # noqa: E402

torch.manual_seed(1234)
Bn, T, K, nb = 8, 16000, 10, 65536
dev = "cuda"
y = (torch.randint(0, nb, (Bn, T), device=dev).float() / (nb - 1) * 2 - 1)
raw = torch.randn(Bn, T, 3 * K, device=dev)
raw[..., K:2 * K] = y.unsqueeze(-1) + 0.1 * torch.randn(Bn, T, K, device=dev)
raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
r = raw.clone().requires_grad_(True)
lik = blvm_b200.DiscretizedLogisticMixtureDense(3, 1, K, nb)
lp = lik.log_prob(y.unsqueeze(-1), blvm_b200.DMoLParams(r, K, 1, -7.0))
lp.sum().backward()
ref_lp, ref_g = O.dmol_value_and_grad(y.cpu().numpy().reshape(-1).astype(np.float64),
                                      raw.cpu().numpy().reshape(-1, 3 * K).astype(np.float64), K, 1, nb)
ours_lp = lp.detach().cpu().numpy().reshape(-1).astype(np.float64)
ours_g = r.grad.cpu().numpy().reshape(-1, 3 * K).astype(np.float64)
e = np.abs(ours_lp - ref_lp) / (1e-5 * np.abs(ref_lp) + 1e-6)
print(f"lp worst err/tol {e.max():.3f}; max rel {np.max(np.abs(ours_lp - ref_lp) / np.abs(ref_lp)):.2e}")
for gi, name in enumerate(("logits", "locs", "log_scales")):
    o, rr = ours_g[:, gi * K:(gi + 1) * K], ref_g[:, gi * K:(gi + 1) * K]
    gmax = np.abs(rr).max(-1, keepdims=True)
    err = np.abs(o - rr)
    ratio = err / (1e-5 * np.abs(rr) + 1e-5 * gmax + 1e-6)
    i, j = np.unravel_index(ratio.argmax(), ratio.shape)
    print(f"{name:10s} worst err/tol {ratio.max():.3f} at sample {i} comp {j}: ours {o[i, j]:.8g} ref {rr[i, j]:.8g} "
          f"gmax {gmax[i, 0]:.4g}; err/gmax pctl 99.9 {np.percentile(err / (gmax + 1e-30), 99.9):.2e} max {np.max(err / (gmax + 1e-30)):.2e}")
    print("   row ref:", np.array2string(rr[i], precision=4))
    print("   row our:", np.array2string(o[i], precision=4))

# KL fused detail
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden  # noqa: E402
g = load_golden("kl_free_nats")
Bq, Tz, Z = g["mu_q"].shape
x_sl = torch.tensor([Tz, Tz - 4, 3])
for i, fn in enumerate(g["free_nats"]):
    ins = [torch.as_tensor(g[n]).cuda().requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
    out = blvm_b200.fused_elbo(None, None, x_sl, [blvm_b200.KLLevel(*ins, stride=1)], beta=0.7, free_nats=float(fn), num_bins=2)
    out.loss.backward()
    m = O.sequence_mask(x_sl.numpy(), max_len=Tz)[..., None].astype(np.float64)
    kl, kl_fn, grads = O.kl_value_and_grad(*[g[n].astype(np.float64) for n in ("mu_q", "sd_q", "mu_p", "sd_p")],
                                           free_nats=float(fn), gout=m * (0.7 / float(x_sl.sum())))
    print(f"fn={fn:.4g} kl rows rel {np.max(np.abs(out.kl.cpu().numpy() - (kl * m).sum((1, 2))) / np.abs((kl * m).sum((1, 2)))):.2e} "
          f"klfn rows rel {np.max(np.abs(out.kl_fn.cpu().numpy() - (kl_fn * m).sum((1, 2))) / np.abs((kl_fn * m).sum((1, 2)))):.2e} "
          f"loss {out.loss.item():.10g} ref {0.7 * (kl_fn * m).sum() / float(x_sl.sum()):.10g}")
    for t, rg, nme in zip(ins, grads, ("mu_q", "sd_q", "mu_p", "sd_p")):
        d = np.abs(t.grad.cpu().numpy() - rg)
        tol = 1e-5 * np.abs(rg) + 1e-7 * np.abs(rg).max()
        k = np.unravel_index((d / tol).argmax(), d.shape)
        print(f"   {nme}: worst err/tol {np.max(d / tol):.3f} at {k} ours {t.grad.cpu().numpy()[k]:.8g} ref {rg[k]:.8g} max|ref| {np.abs(rg).max():.4g}")

#!/usr/bin/env python
"""Generate the golden input/output fixtures under tests/golden/*.npz from the REFERENCE itself.

Run in the build container only (the reference tree does not exist on the GPU box):

    BLVM_DATA_ROOT_DIRECTORY=/tmp/blvmdata \
    PYTHONPATH=/root/reference:tests/golden/_ref_shims python tests/golden/make_golden.py

Everything here calls the unmodified functions of `/root/reference/blvm` (file:line cited per case) on
seeded CPU inputs, in fp32 (the reference's native arithmetic) and in fp64 (same functions, `.double()`
inputs; they are dtype-generic) and stores inputs + outputs + autograd gradients as small `.npz` files.
The reference's own tests pin nothing on this path, so these fixtures are the pin for
both `oracle/` and the CUDA path.
"""
import math
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

os.environ.setdefault("BLVM_DATA_ROOT_DIRECTORY", "/tmp/blvmdata")
os.makedirs(os.environ["BLVM_DATA_ROOT_DIRECTORY"], exist_ok=True)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref_shims"))
sys.path.insert(0, "/root/reference")

from blvm.utils.log_likelihoods import discretized_logistic_ll, discretized_logistic_mixture_ll  # noqa: E402
from blvm.utils.variational import discount_free_nats, kl_divergence_gaussian  # noqa: E402
from blvm.utils.operations import sequence_mask  # noqa: E402
from blvm.modules.distributions import DiscretizedLogisticMixtureDense, DiscretizedLogisticDense  # noqa: E402
from blvm.data.transforms import Quantize  # noqa: E402
from blvm.evaluation.metrics import BitsPerDimMetric, LLMetric, KLMetric, LossMetric  # noqa: E402
import blvm.models  # noqa: E402,F401
import importlib  # noqa: E402

ref_vrnn = importlib.import_module("blvm.models.vrnn")
ref_srnn = importlib.import_module("blvm.models.srnn")
ref_stcn = importlib.import_module("blvm.models.stcn.stcn")
ref_cwvae = importlib.import_module("blvm.models.clockwork_vae.clockwork_vae")
ref_wavenet = importlib.import_module("blvm.models.wavenet.wavenet")

torch.set_num_threads(1)  # bit-reproducible reductions
LOG_EPS = -7.0


def f32(x):
    return np.float32(x)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ----------------------------------------------------------------------------------------------------------------
# (i) DMoL: blvm/modules/distributions.py:381-387 (split + clamp) -> blvm/utils/log_likelihoods.py:170-231
# ----------------------------------------------------------------------------------------------------------------
def ref_dmol_from_raw(y, raw, K, D, nb, gout, dtype):
    """y (N, D), raw (N, K(2D+1)) -> per-sample log-prob (N,) and d(sum gout*lp)/d raw, via the reference."""
    y = y.to(dtype)
    raw = raw.to(dtype).clone().requires_grad_(True)
    logit_probs = raw[..., :K]
    locs_log_scales = raw[..., K:].view(*raw.shape[:-1], D, 2 * K)
    locs, log_scales = locs_log_scales.chunk(2, dim=-1)
    log_scales = log_scales.clamp(min=LOG_EPS)
    lp = discretized_logistic_mixture_ll(y, logit_probs, locs, log_scales, num_bins=nb)
    (lp * gout.to(dtype)).sum().backward()
    return lp.detach(), raw.grad.detach()


def special_y(nb, n, gen):
    lo = f32(2 / nb - 1)
    hi = f32(1 - 2 / nb)
    specials = [
        -1.0, 1.0, 0.0,
        lo, np.nextafter(lo, f32(-2)), np.nextafter(lo, f32(2)),
        hi, np.nextafter(hi, f32(-2)), np.nextafter(hi, f32(2)),
    ]
    y = torch.randint(0, nb, (n,), generator=gen).float() / (nb - 1) * 2 - 1  # the reference's rescaled grid
    y16 = torch.randint(-32768, 32768, (n,), generator=gen).float() / 32768  # int16 PCM grid
    y = torch.where(torch.rand(n, generator=gen) < 0.5, y, y16)
    k = min(len(specials), n)
    y[:k] = torch.tensor(np.array(specials[:k], dtype=np.float32))
    # a few more exact edge hits spread through the batch
    y[k: k + 4] = torch.tensor([-1.0, 1.0, float(lo), float(hi)])[: max(0, min(4, n - k))]
    return y.clamp(-1, 1)


def make_dmol_case(K, nb, N, seed, D=1):
    gen = torch.Generator().manual_seed(seed)
    P = K * (2 * D + 1)
    y = torch.stack([special_y(nb, N, gen) for _ in range(D)], dim=-1)  # (N, D)
    raw = torch.randn(N, P, generator=gen)
    raw[:, :K] *= 1.5
    lls = raw[:, K:].view(N, D, 2 * K)
    # locs: half of them near the target (peaky regime), half anywhere
    near = torch.rand(N, D, K, generator=gen) < 0.5
    lls[..., :K] = torch.where(
        near, y.unsqueeze(-1) + 0.02 * torch.randn(N, D, K, generator=gen), torch.rand(N, D, K, generator=gen) * 2.4 - 1.2
    )
    lls[..., K:] = torch.randn(N, D, K, generator=gen) * 2 - 4  # straddles the -7 clamp and the delta threshold
    # forced rows (only touch component 0..min(K,3) so that every K gets them)
    r = 16
    h = 1.0 / (nb - 1)
    if N >= 64:
        lls[r + 0, 0, K + 0] = -7.0          # clamp tie: gradient passes (SURVEY §7 tie rules)
        lls[r + 1, 0, K + 0] = -7.5          # clamped: zero gradient
        lls[r + 2, 0, K + 0] = float(np.nextafter(f32(-7), f32(-8)))
        lls[r + 3, 0, K + 0] = 5.0           # very wide
        y[r + 4, 0] = 0.9; lls[r + 4, 0, 0] = -0.9; lls[r + 4, 0, K] = -6.0   # |a| >> 20 (softplus threshold)
        y[r + 5, 0] = -0.9; lls[r + 5, 0, 0] = 0.9; lls[r + 5, 0, K] = -6.5
        raw[r + 6, :K] = -50.0; raw[r + 6, 0] = 50.0   # one component owns the mixture
        raw[r + 7, :K] = 0.0                              # uniform mixture
        # delta straddling 1e-5 at m ~ 0: delta ~ 0.5*h*inv  => inv ~ 2e-5/h
        inv0 = 2e-5 / h
        for j in range(24):
            lls[r + 8 + j, 0, 0] = y[r + 8 + j, 0]
            lls[r + 8 + j, 0, K] = -math.log(inv0) + (j - 12) * 2e-3
            if K > 1:
                raw[r + 8 + j, 1:K] = -30.0  # let component 0 dominate so the branch is visible in lp
        # exact-centre hit and huge logits
        lls[r + 32, 0, 0] = y[r + 32, 0]
        raw[r + 33, :K] = 80.0 * torch.sign(torch.randn(K, generator=gen))
    y = y.clamp(-1, 1)
    gout = torch.randn(N, generator=gen)
    gout[:4] = 1.0
    lp32, g32 = ref_dmol_from_raw(y, raw, K, D, nb, gout, torch.float32)
    lp64, g64 = ref_dmol_from_raw(y, raw, K, D, nb, gout, torch.float64)
    return dict(
        y=y.numpy(), raw=raw.numpy(), gout=gout.numpy(), K=K, D=D, num_bins=nb, log_epsilon=LOG_EPS,
        lp32=lp32.numpy(), graw32=g32.numpy(), lp64=lp64.numpy(), graw64=g64.numpy(),
    )


def make_dmol():
    for K in (1, 2, 10, 30):
        for nb in (256, 65536):
            save(f"dmol_K{K}_nb{nb}", **make_dmol_case(K, nb, 384, seed=1000 + 7 * K + (nb == 256)))
    save("dmol_K10_nb65536_D2", **make_dmol_case(10, 65536, 192, seed=77, D=2))
    save("dmol_K5_nb255", **make_dmol_case(5, 255, 192, seed=78))  # non power-of-two bins: fp32-rounded constants


# ----------------------------------------------------------------------------------------------------------------
# (i-b) DL: blvm/modules/distributions.py:303-307 -> blvm/utils/log_likelihoods.py:98-166
# ----------------------------------------------------------------------------------------------------------------
def make_dl():
    for nb in (256, 65536):
        gen = torch.Generator().manual_seed(2000 + nb)
        N = 384
        y = special_y(nb, N, gen)
        raw = torch.randn(N, 2, generator=gen)
        raw[:, 0] = torch.where(torch.rand(N, generator=gen) < 0.5, y + 0.02 * torch.randn(N, generator=gen), raw[:, 0])
        raw[:, 1] = raw[:, 1] * 2 - 4
        raw[16, 1] = -7.0
        raw[17, 1] = -7.5
        raw[18, 1] = 5.0
        gout = torch.randn(N, generator=gen)
        out = {}
        for tag, dt in (("32", torch.float32), ("64", torch.float64)):
            r = raw.to(dt).clone().requires_grad_(True)
            mu, ls = r.chunk(2, dim=-1)
            ls = ls.clamp(min=LOG_EPS)
            lp = discretized_logistic_ll(y.to(dt).unsqueeze(-1), mu, ls, num_bins=nb, reduce_dim=None)  # (N, 1)
            (lp.squeeze(-1) * gout.to(dt)).sum().backward()
            out["lp" + tag] = lp.detach().squeeze(-1).numpy()
            out["graw" + tag] = r.grad.numpy()
        save(f"dl_nb{nb}", y=y.numpy(), raw=raw.numpy(), gout=gout.numpy(), num_bins=nb, log_epsilon=LOG_EPS, **out)


# ----------------------------------------------------------------------------------------------------------------
# (ii) KL + free nats: blvm/utils/variational.py:67-70, :86-122
# ----------------------------------------------------------------------------------------------------------------
def make_kl():
    gen = torch.Generator().manual_seed(3000)
    B, Tz, Z = 3, 11, 16
    mu_q = torch.randn(B, Tz, Z, generator=gen)
    mu_p = torch.randn(B, Tz, Z, generator=gen)
    sd_q = torch.nn.functional.softplus(torch.randn(B, Tz, Z, generator=gen)) + 1e-6
    sd_p = torch.nn.functional.softplus(torch.randn(B, Tz, Z, generator=gen)) + 1e-6
    # identical distributions => kl == 0 exactly
    mu_q[0, 0, :4] = mu_p[0, 0, :4]
    sd_q[0, 0, :4] = sd_p[0, 0, :4]
    sd_q[0, 1, 0] = 1e-4
    sd_p[0, 1, 1] = 1e-3
    sd_p[0, 1, 2] = 30.0
    gout = torch.randn(B, Tz, Z, generator=gen)
    kl32_plain = kl_divergence_gaussian(mu_q, sd_q, mu_p, sd_p)
    tie_value = float(kl32_plain[1, 2, 3])  # fn/Z == this element bit-exactly (Z is a power of two)
    free_nats = [0.0, 0.0625, 4.0, tie_value * Z]
    out = dict(mu_q=mu_q.numpy(), sd_q=sd_q.numpy(), mu_p=mu_p.numpy(), sd_p=sd_p.numpy(), gout=gout.numpy(),
               free_nats=np.array(free_nats, dtype=np.float64), tie_index=np.array([1, 2, 3]))
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        for i, fn in enumerate(free_nats):
            ins = [t.to(dt).clone().requires_grad_(True) for t in (mu_q, sd_q, mu_p, sd_p)]
            kl = kl_divergence_gaussian(*ins)
            kl_fn = discount_free_nats(kl, fn, shared_dims=-1)
            (kl_fn * gout.to(dt)).sum().backward()
            if i == 0:
                out["kl" + tag] = kl.detach().numpy()
            out[f"klfn{tag}_{i}"] = kl_fn.detach().numpy()
            for nme, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), ins):
                out[f"g_{nme}{tag}_{i}"] = t.grad.numpy()
    save("kl_free_nats", **out)


# ----------------------------------------------------------------------------------------------------------------
# (iii) the per-model ELBO reducers, called unbound on a stub `self`
# ----------------------------------------------------------------------------------------------------------------
def _kl_inputs(gen, B, Tz, Z):
    mu_q = torch.randn(B, Tz, Z, generator=gen)
    mu_p = torch.randn(B, Tz, Z, generator=gen)
    sd_q = torch.nn.functional.softplus(torch.randn(B, Tz, Z, generator=gen)) + 1e-3
    sd_p = torch.nn.functional.softplus(torch.randn(B, Tz, Z, generator=gen)) + 1e-3
    return [mu_q, sd_q, mu_p, sd_p]


def _dmol_inputs(gen, B, T, K, nb):
    y = torch.stack([special_y(nb, T, gen) for _ in range(B)])  # (B, T)
    raw = torch.randn(B, T, 3 * K, generator=gen)
    raw[..., K:2 * K] = torch.where(torch.rand(B, T, K, generator=gen) < 0.5,
                                    y.unsqueeze(-1) + 0.02 * torch.randn(B, T, K, generator=gen),
                                    torch.rand(B, T, K, generator=gen) * 2.4 - 1.2)
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    return y, raw


def _likelihood(K, nb):
    lk = DiscretizedLogisticMixtureDense(x_dim=3 * K, y_dim=1, num_mix=K, num_bins=nb)
    return lk


def _params_from_raw(raw, K):
    """What DiscretizedLogisticMixtureDense.forward does after the Linear (distributions.py:383-387)."""
    logit_probs = raw[..., :K]
    lls = raw[..., K:].view(*raw.shape[:-1], 1, 2 * K)
    locs, log_scales = lls.chunk(2, dim=-1)
    return logit_probs, locs, log_scales.clamp(min=LOG_EPS)


def _np(t):
    return t.detach().numpy()


def make_elbo_models():
    K, nb = 10, 65536
    beta, free_nats = 0.5, 0.0625
    # ---- VRNN (vrnn.py:255-279) and SRNN (srnn.py:137-160): float64 masks, stride S, one KL tensor --------------
    for model_name, fn in (("vrnn", ref_vrnn.VRNN.compute_elbo), ("srnn", ref_srnn.SRNN.compute_elbo)):
        for variant, (bt, fnats) in {"a": (beta, free_nats), "b": (1.0, 0.0)}.items():
            gen = torch.Generator().manual_seed(4000 + len(model_name) + ord(variant))
            B, T, S, Z = 4, 96, 8, 16
            x_sl = torch.tensor([96, 50, 33, 8])
            y, raw = _dmol_inputs(gen, B, T, K, nb)
            kl_in = _kl_inputs(gen, B, T // S, Z)
            out = dict(y=_np(y), raw=_np(raw), x_sl=x_sl.numpy(), stride=S, beta=bt, free_nats=fnats, K=K, num_bins=nb,
                       **{n: _np(t) for n, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), kl_in)})
            for tag, dt in (("32", torch.float32), ("64", torch.float64)):
                r = raw.to(dt).clone().requires_grad_(True)
                ins = [t.to(dt).clone().requires_grad_(True) for t in kl_in]
                self = SimpleNamespace(likelihood=_likelihood(K, nb))
                kld = kl_divergence_gaussian(*ins)
                loss, elbo, logp, kl, seq_mask = fn(self, y.to(dt).unsqueeze(-1), _params_from_raw(r, K), kld, x_sl, S, bt, fnats)
                loss.backward()
                out.update({f"loss{tag}": _np(loss), f"elbo{tag}": _np(elbo), f"logp{tag}": _np(logp), f"kl{tag}": _np(kl),
                            f"graw{tag}": _np(r.grad), f"dtype{tag}": str(elbo.dtype), f"mask_dtype{tag}": str(seq_mask.dtype)})
                for nme, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), ins):
                    out[f"g_{nme}{tag}"] = _np(t.grad)
                if tag == "32":
                    out["bpd32"] = np.float64(BitsPerDimMetric(elbo, reduce_by=x_sl).value)
            save(f"elbo_{model_name}_{variant}", **out)

    # ---- Clockwork-VAE (clockwork_vae.py:132-161, masks :231-240): bool masks, 3 levels, scaled free nats ---------
    gen = torch.Generator().manual_seed(4100)
    B, T = 3, 128
    strides = [4, 2, 2]
    ostr = [4, 8, 16]
    Zs = [8, 4, 4]
    x_sl = torch.tensor([128, 77, 20])
    y, raw = _dmol_inputs(gen, B, T, K, nb)
    kl_ins = [_kl_inputs(gen, B, T // s, z) for s, z in zip(ostr, Zs)]
    out = dict(y=_np(y), raw=_np(raw), x_sl=x_sl.numpy(), overall_strides=np.array(ostr), beta=beta, free_nats=free_nats,
               K=K, num_bins=nb, num_levels=3)
    for l, ins in enumerate(kl_ins):
        out.update({f"{n}_{l}": _np(t) for n, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), ins)})
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        r = raw.to(dt).clone().requires_grad_(True)
        lv = [[t.to(dt).clone().requires_grad_(True) for t in ins] for ins in kl_ins]
        self = SimpleNamespace(likelihood=_likelihood(K, nb), num_levels=3, overall_strides=ostr)
        seq_mask = sequence_mask(x_sl, max_len=T)
        level_masks = [sequence_mask((x_sl / s).ceil().int(), max_len=T // s) for s in ostr]  # cwvae :237-238
        klds = [kl_divergence_gaussian(*ins) for ins in lv]
        loss, elbo, logp, kld, kld_l = ref_cwvae.CWVAE.compute_elbo(
            self, y.to(dt).unsqueeze(-1), seq_mask, level_masks, x_sl, _params_from_raw(r, K), klds, beta, free_nats)
        loss.backward()
        out.update({f"loss{tag}": _np(loss), f"elbo{tag}": _np(elbo), f"logp{tag}": _np(logp), f"kl{tag}": _np(kld),
                    f"graw{tag}": _np(r.grad), f"dtype{tag}": str(elbo.dtype)})
        for l in range(3):
            out[f"kl_l{l}_{tag}"] = _np(kld_l[l])
            for nme, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), lv[l]):
                out[f"g_{nme}_{l}_{tag}"] = _np(t.grad)
    save("elbo_cwvae", **out)

    # ---- STCN (stcn.py:256-297): bool mask, n_latents levels at the same stride, mask-fn-mask -------------------
    gen = torch.Generator().manual_seed(4200)
    B, T, S = 3, 64, 1
    n_lat = 3
    Zs = [8, 4, 2]
    x_sl = torch.tensor([64, 41, 7])
    y, raw = _dmol_inputs(gen, B, T, K, nb)
    kl_ins = [_kl_inputs(gen, B, T // S, z) for z in Zs]
    out = dict(y=_np(y), raw=_np(raw), x_sl=x_sl.numpy(), n_stack_frames=S, beta=beta, free_nats=free_nats, K=K,
               num_bins=nb, n_latents=n_lat)
    for l, ins in enumerate(kl_ins):
        out.update({f"{n}_{l}": _np(t) for n, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), ins)})
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        r = raw.to(dt).clone().requires_grad_(True)
        lv = [[t.to(dt).clone().requires_grad_(True) for t in ins] for ins in kl_ins]
        self = SimpleNamespace(likelihood_module=_likelihood(K, nb), n_stack_frames=S, top_down=True, n_latents=n_lat)
        mu_q, sd_q, mu_p, sd_p = ([ins[i] for ins in lv] for i in range(4))
        loss, elbo, logp, kld, klds = ref_stcn.STCN.compute_loss(
            self, y.to(dt).unsqueeze(-1), x_sl, _params_from_raw(r, K), mu_p, sd_p, mu_q, sd_q, None, free_nats, beta)
        loss.backward()
        out.update({f"loss{tag}": _np(loss), f"elbo{tag}": _np(elbo), f"logp{tag}": _np(logp), f"kl{tag}": _np(kld),
                    f"graw{tag}": _np(r.grad), f"dtype{tag}": str(elbo.dtype)})
        for l in range(n_lat):
            out[f"kl_l{l}_{tag}"] = _np(klds[l])
            for nme, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), lv[l]):
                out[f"g_{nme}_{l}_{tag}"] = _np(t.grad)
    save("elbo_stcn", **out)

    # ---- WaveNet (wavenet.py:128-146): bool mask with max_len, nansum ---------------------------------------------
    gen = torch.Generator().manual_seed(4300)
    B, T = 4, 80
    x_sl = torch.tensor([80, 64, 31, 1])
    y, raw = _dmol_inputs(gen, B, T, K, nb)
    out = dict(y=_np(y), raw=_np(raw), x_sl=x_sl.numpy(), K=K, num_bins=nb)
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        r = raw.to(dt).clone().requires_grad_(True)
        self = SimpleNamespace(likelihood=_likelihood(K, nb))
        loss, logp, logp_twise = ref_wavenet.WaveNet.compute_loss(self, y.to(dt).unsqueeze(-1), x_sl, _params_from_raw(r, K))
        loss.backward()
        out.update({f"loss{tag}": _np(loss), f"logp{tag}": _np(logp), f"logp_twise{tag}": _np(logp_twise),
                    f"graw{tag}": _np(r.grad), f"dtype{tag}": str(logp.dtype)})
        if tag == "32":
            out["bpd32"] = np.float64(BitsPerDimMetric(logp, reduce_by=x_sl).value)
    save("elbo_wavenet", **out)


# ----------------------------------------------------------------------------------------------------------------
# (iv) Quantize: blvm/data/transforms.py:216-260 -- integer bin indices, bit-exact
# ----------------------------------------------------------------------------------------------------------------
def make_quantize():
    pcm = torch.arange(-32768, 32768, dtype=torch.float32) / 32768  # every int16 PCM value
    gen = torch.Generator().manual_seed(5000)
    rnd = torch.rand(4096, generator=gen) * 2 - 1
    out = dict(rnd=rnd.numpy())
    for bits in (8, 16):
        q = Quantize(bits=bits)
        grid = torch.arange(0, 2 ** bits, dtype=torch.float32) / (2 ** bits - 1) * 2 - 1
        out[f"boundaries_{bits}"] = q.boundaries.numpy()
        out[f"pcm_idx_{bits}"] = q(pcm).numpy().astype(np.int32)
        out[f"rnd_idx_{bits}"] = q(rnd).numpy().astype(np.int32)
        out[f"grid_idx_{bits}"] = q(grid).numpy().astype(np.int32)
        out[f"bnd_idx_{bits}"] = q(q.boundaries).numpy().astype(np.int32)  # exact boundary hits (right=False)
    save("quantize", **out)


# ----------------------------------------------------------------------------------------------------------------
# (v) sibling likelihoods: GMM (distributions.py:153-204 + log_likelihoods.py:42-60), Gaussian (:17-39), MC KL
# ----------------------------------------------------------------------------------------------------------------
def make_gmm():
    from blvm.modules.distributions import DiagonalGaussianMixtureDense
    from blvm.utils.log_likelihoods import gaussian_ll, gaussian_mixture_ll
    from blvm.utils.variational import kl_divergence_gaussian_mc
    for K in (1, 5, 10, 20, 7):
        gen = torch.Generator().manual_seed(6000 + K)
        N, D = 320, 1
        mod = DiagonalGaussianMixtureDense(x_dim=3 * K, y_dim=D, num_mix=K, initial_sd=1, epsilon=1e-4)
        y = torch.rand(N, D, generator=gen) * 2 - 1
        raw = torch.randn(N, 3 * K, generator=gen)
        raw[:, K:2 * K] = y + 0.3 * torch.randn(N, K, generator=gen)
        raw[:, 2 * K:] = raw[:, 2 * K:] * 4 - 4            # softplus regime from ~exp(-20) to linear
        raw[0, 2 * K] = 40.0                               # beta*x > threshold: linear branch
        raw[1, 2 * K] = -60.0                              # sd == epsilon
        gout = torch.randn(N, generator=gen)
        out = dict(y=y.numpy(), raw=raw.numpy(), gout=gout.numpy(), K=K, D=D, beta=math.log(2) / 1.0, sd_add=1e-4)
        for tag, dt in (("32", torch.float32), ("64", torch.float64)):
            r = raw.to(dt).clone().requires_grad_(True)
            logits = r[..., :K]
            mls = r[..., K:].view(N, D, 2 * K)
            mu, log_sd = mls.chunk(2, dim=-1)
            sd = mod.sd_activation(log_sd)                  # distributions.py:203
            lp = gaussian_mixture_ll(y.to(dt), logits, mu, sd, epsilon=0)
            (lp * gout.to(dt)).sum().backward()
            out["lp" + tag] = lp.detach().numpy()
            out["graw" + tag] = r.grad.numpy()
            if tag == "64":
                out["sd64"] = sd.detach().numpy()
        save(f"gmm_K{K}", **out)
    # elementwise gaussian_ll (epsilon = 0 as the modules call it, and 1e-2: clamp under no_grad) and the MC KL
    gen = torch.Generator().manual_seed(6100)
    shape = (5, 9, 4)
    y = torch.randn(shape, generator=gen)
    mu_q, mu_p = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)
    sd_q = torch.nn.functional.softplus(torch.randn(shape, generator=gen)) + 1e-3
    sd_p = torch.nn.functional.softplus(torch.randn(shape, generator=gen)) + 1e-3
    sd_q[0, 0, 0] = 1e-3
    gout = torch.randn(shape, generator=gen)
    out = dict(y=y.numpy(), mu_q=mu_q.numpy(), sd_q=sd_q.numpy(), mu_p=mu_p.numpy(), sd_p=sd_p.numpy(), gout=gout.numpy())
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        for eps in (0.0, 1e-2):
            m = mu_q.to(dt).clone().requires_grad_(True)
            s_ = sd_q.to(dt).clone().requires_grad_(True)
            lp = gaussian_ll(y.to(dt), m, s_, epsilon=eps, reduce_dim=None)
            (lp * gout.to(dt)).sum().backward()
            e = "0" if eps == 0 else "1"
            out[f"lp{tag}_{e}"] = lp.detach().numpy()
            out[f"g_mu{tag}_{e}"] = m.grad.numpy()
            out[f"g_sd{tag}_{e}"] = s_.grad.numpy() if s_.grad is not None else np.zeros(shape)
        ins = [t.to(dt).clone().requires_grad_(True) for t in (mu_q, sd_q, mu_p, sd_p)]
        kl = kl_divergence_gaussian_mc(*ins, y.to(dt))
        (kl * gout.to(dt)).sum().backward()
        out["klmc" + tag] = kl.detach().numpy()
        for nme, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), ins):
            out[f"klmc_g_{nme}{tag}"] = t.grad.numpy()
    save("gaussian_ll", **out)

def make_stcn_bottom_up():
    """STCN with top_down=False (stcn.py:286-288): the KL of every level is the Monte-Carlo estimate log q(z) - log p(z)
    (variational.py:73-83), which can be negative; mask -> free nats -> mask as in the analytic variant."""
    K, nb = 10, 65536
    beta, free_nats = 0.5, 0.0625
    gen = torch.Generator().manual_seed(4250)
    B, T, S = 3, 64, 1
    Zs = [8, 4]
    x_sl = torch.tensor([64, 33, 5])
    y, raw = _dmol_inputs(gen, B, T, K, nb)
    kl_ins = [_kl_inputs(gen, B, T // S, z) for z in Zs]
    zs = [ins[0] + ins[1] * torch.randn(ins[0].shape, generator=gen) for ins in kl_ins]       # z ~ q
    out = dict(y=_np(y), raw=_np(raw), x_sl=x_sl.numpy(), n_stack_frames=S, beta=beta, free_nats=free_nats, K=K,
               num_bins=nb, n_latents=len(Zs))
    for l, ins in enumerate(kl_ins):
        out.update({f"{n}_{l}": _np(t) for n, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), ins)})
        out[f"z_{l}"] = _np(zs[l])
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        r = raw.to(dt).clone().requires_grad_(True)
        lv = [[t.to(dt).clone().requires_grad_(True) for t in ins] for ins in kl_ins]
        self = SimpleNamespace(likelihood_module=_likelihood(K, nb), n_stack_frames=S, top_down=False, n_latents=len(Zs))
        mu_q, sd_q, mu_p, sd_p = ([ins[i] for ins in lv] for i in range(4))
        zz = [z.to(dt).clone().requires_grad_(True) for z in zs]       # in the model z = rsample(q) carries gradient
        loss, elbo, logp, kld, klds = ref_stcn.STCN.compute_loss(
            self, y.to(dt).unsqueeze(-1), x_sl, _params_from_raw(r, K), mu_p, sd_p, mu_q, sd_q, zz, free_nats, beta)
        loss.backward()
        for l in range(len(Zs)):
            out[f"g_z_{l}_{tag}"] = _np(zz[l].grad)
        out.update({f"loss{tag}": _np(loss), f"elbo{tag}": _np(elbo), f"logp{tag}": _np(logp), f"kl{tag}": _np(kld),
                    f"graw{tag}": _np(r.grad), f"dtype{tag}": str(elbo.dtype)})
        for l in range(len(Zs)):
            out[f"kl_l{l}_{tag}"] = _np(klds[l])
            for nme, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), lv[l]):
                out[f"g_{nme}_{l}_{tag}"] = _np(t.grad)
    save("elbo_stcn_bottom_up", **out)


if __name__ == "__main__":
    if "--stcn-bu-only" in sys.argv:
        make_stcn_bottom_up()
        sys.exit(0)
    if "--gmm-only" in sys.argv:
        make_gmm()
        sys.exit(0)
    make_dmol()
    make_dl()
    make_kl()
    make_elbo_models()
    make_quantize()
    make_gmm()
    make_stcn_bottom_up()

"""GPU parity: the CUDA path (through the Python boundary -> ctypes -> C ABI -> sm_100a kernels) against the golden
vectors generated from the reference and against the numpy oracle.  Run on the B200 box: pytest -m gpu."""
import math
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import load_golden
from parity import RTOL, assert_grads_close, assert_sums_close, assert_values_close, gmm_row_factor

pytestmark = pytest.mark.gpu

DMOL_CASES = ["dmol_K1_nb256", "dmol_K1_nb65536", "dmol_K2_nb256", "dmol_K2_nb65536", "dmol_K10_nb256",
              "dmol_K10_nb65536", "dmol_K30_nb256", "dmol_K30_nb65536", "dmol_K5_nb255"]


@pytest.fixture(scope="module")
def B():
    import blvm_b200
    assert torch.cuda.is_available()
    return blvm_b200


@pytest.fixture(scope="module")
def O():
    from oracle import blvm_oracle
    return blvm_oracle


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(dtype).cuda()


def _edge_rows(g, O):
    """Rows where the y-edge predicate evaluated in fp64 (golden *64) differs from fp32 (the reference's native
    arithmetic; only possible for non power-of-two num_bins): there the fp32 run is the truth."""
    nb = int(g["num_bins"])
    y32 = g["y"].astype(np.float32)
    lo32, hi32 = np.float32(2 / nb - 1), np.float32(1 - 2 / nb)
    e32 = np.where(y32 > hi32, 2, np.where(y32 < lo32, 1, 0))
    y64 = g["y"].astype(np.float64)
    e64 = np.where(y64 > 1 - 2 / nb, 2, np.where(y64 < 2 / nb - 1, 1, 0))
    return (e32 != e64).any(-1)


@pytest.mark.parametrize("case", DMOL_CASES)
def test_dmol_module_golden(case, B, O):
    g = load_golden(case)
    K, D, nb = int(g["K"]), int(g["D"]), int(g["num_bins"])
    lik = B.DiscretizedLogisticMixtureDense(3 * K, D, num_mix=K, num_bins=nb)
    raw = cu(g["raw"]).requires_grad_(True)
    params = B.DMoLParams(raw, K, D, lik.log_epsilon)
    lp = lik.log_prob(cu(g["y"]), params)
    assert lp.shape == (raw.shape[0],) and lp.dtype == torch.float32
    (lp * cu(g["gout"])).sum().backward()
    B.check_input_range()
    flip = _edge_rows(g, O)
    # knife-edge rows: a component whose cdf_delta sits within 1e-5 relative of the 1e-5 threshold (log_likelihoods.py:222).
    # The reference's two arms differ there by log(nb/(nb-1)) (bin width 2/nb vs 2/(nb-1)), so which side an
    # implementation lands on is decided by its last ulps; checked separately below.
    _, delta64 = O.dmol_branches(g["y"].astype(np.float64), g["raw"].astype(np.float64), K, D, nb)
    knife = (np.abs(delta64 / 1e-5 - 1) < 1e-5).reshape(len(flip), -1).any(-1)
    ok = ~flip & ~knife
    if knife.any():
        d = np.abs(lp.detach().cpu().numpy()[knife] - g["lp64"][knife])
        assert (d <= math.log(nb / (nb - 1)) + 1e-4).all()
    assert_values_close(lp.detach().cpu().numpy()[ok], g["lp64"][ok], "log-prob vs reference fp64")
    assert_grads_close(raw.grad.cpu().numpy()[ok], g["graw64"][ok], K, np.abs(g["gout"])[ok], "d/d raw vs reference fp64")
    if flip.any():  # fp32 predicates are the reference's: compare those rows with its fp32 run, loosely (see parity.py)
        np.testing.assert_allclose(lp.detach().cpu().numpy()[flip], g["lp32"][flip], rtol=2e-3)
    # where the reference's fp32 run is self-consistent with its fp64 run, we also match it directly
    good = ok & (np.abs(g["lp32"] - g["lp64"]) <= 1e-6 * np.abs(g["lp64"]))
    assert_values_close(lp.detach().cpu().numpy()[good], g["lp32"][good], "log-prob vs reference fp32", rtol=1.1e-5)
    # and we are never further from the fp64 truth than the fp32 reference is (plus the tolerance)
    ours_err = np.abs(lp.detach().cpu().numpy().astype(np.float64) - g["lp64"])[ok]
    ref_err = np.abs(g["lp32"].astype(np.float64) - g["lp64"])[ok]
    assert (ours_err <= ref_err + RTOL * np.abs(g["lp64"][ok]) + 1e-6).all()


def test_dmol_branch_coverage(O):
    """The golden batches exercise all four branches of log_likelihoods.py:221-227 in both bin widths."""
    for case in ("dmol_K10_nb256", "dmol_K10_nb65536"):
        g = load_golden(case)
        br, _ = O.dmol_branches(g["y"].astype(np.float64), g["raw"].astype(np.float64), 10, 1, int(g["num_bins"]))
        assert set(np.unique(br)) == {0, 1, 2, 3}


def test_dmol_generic_kernel_D2(B):
    g = load_golden("dmol_K10_nb65536_D2")
    K, D, nb = 10, 2, 65536
    lik = B.DiscretizedLogisticMixtureDense(3, D, num_mix=K, num_bins=nb)
    raw = cu(g["raw"]).requires_grad_(True)
    lp = lik.log_prob(cu(g["y"]), B.DMoLParams(raw, K, D, -7.0))
    (lp * cu(g["gout"])).sum().backward()
    assert_values_close(lp.detach().cpu().numpy(), g["lp64"], "D=2 log-prob")
    assert_grads_close(raw.grad.cpu().numpy(), g["graw64"], K, np.abs(g["gout"]), "D=2 grads")


@pytest.mark.parametrize("K", [7, 10])
def test_dmol_functional_unpacked_and_broadcast(K, B, O):
    """Functional API with separate (already clamped) tensors, a K without a register kernel (7 -> generic kernel), and
    broadcast parameters as in experiments/experiment_distribution_audio.py:122-131."""
    rng = np.random.default_rng(K)
    N, nb = 300, 65536
    y = (rng.integers(0, nb, (N, 1)) / (nb - 1) * 2 - 1).astype(np.float32)
    logits = rng.normal(size=(N, K)).astype(np.float32)
    locs = (y[:, None, :] + 0.05 * rng.normal(size=(N, 1, K))).astype(np.float32)
    ls = np.maximum(rng.normal(size=(N, 1, K)) * 2 - 4, -7).astype(np.float32)
    t = [cu(a).requires_grad_(True) for a in (logits, locs, ls)]
    lp = B.discretized_logistic_mixture_ll(cu(y), *t, num_bins=nb)
    lp.sum().backward()
    ref = O.discretized_logistic_mixture_ll(y.astype(np.float64), logits.astype(np.float64), locs.astype(np.float64),
                                            ls.astype(np.float64), nb)
    assert_values_close(lp.detach().cpu().numpy(), ref, "functional log-prob")
    raw = np.concatenate([logits, locs[:, 0], ls[:, 0]], -1).astype(np.float64)
    _, gref = O.dmol_value_and_grad(y.astype(np.float64), raw, K, 1, nb, log_epsilon=-np.inf)
    ours = np.concatenate([t[0].grad.cpu().numpy(), t[1].grad.cpu().numpy()[:, 0], t[2].grad.cpu().numpy()[:, 0]], -1)
    assert_grads_close(ours, gref, K, np.ones(N), "functional grads")
    # broadcast (K,) / (1, K) parameters against y (B, T, 1)
    yb = cu(y[:64].reshape(4, 16, 1))
    lpb = B.discretized_logistic_mixture_ll(yb, cu(logits[0]), cu(locs[0]), cu(ls[0]), num_bins=nb)
    refb = O.discretized_logistic_mixture_ll(y[:64].reshape(4, 16, 1).astype(np.float64), logits[0].astype(np.float64),
                                             locs[0].astype(np.float64), ls[0].astype(np.float64), nb)
    assert lpb.shape == (4, 16)
    assert_values_close(lpb.cpu().numpy(), refb, "broadcast log-prob")


@pytest.mark.parametrize("nb", [256, 65536])
def test_dl_golden(nb, B):
    g = load_golden(f"dl_nb{nb}")
    lik = B.DiscretizedLogisticDense(4, 1, num_bins=nb)
    raw = cu(g["raw"]).requires_grad_(True)
    lp = lik.log_prob(cu(g["y"]).unsqueeze(-1), B.DLParams(raw, 1, -7.0))
    assert lp.shape == (raw.shape[0], 1)
    (lp.squeeze(-1) * cu(g["gout"])).sum().backward()
    assert_values_close(lp.detach().cpu().numpy()[:, 0], g["lp64"], "DL log-prob")
    assert_grads_close(raw.grad.cpu().numpy(), g["graw64"], 1, np.abs(g["gout"]), "DL grads")
    # functional form with separate tensors
    mu, ls = cu(g["raw"][:, 0]), cu(np.maximum(g["raw"][:, 1], -7.0))
    lp2 = B.discretized_logistic_ll(cu(g["y"]), mu, ls, num_bins=nb, reduce_dim=None)
    assert_values_close(lp2.cpu().numpy(), g["lp64"], "DL functional")


def test_kl_gaussian_golden(B):
    g = load_golden("kl_free_nats")
    ins = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
    kl = B.kl_divergence_gaussian(*ins)
    (kl * cu(g["gout"])).sum().backward()
    assert_values_close(kl.detach().cpu().numpy(), g["kl64"], "KL", atol=1e-6)
    for t, n in zip(ins, ("mu_q", "sd_q", "mu_p", "sd_p")):
        ref = g[f"g_{n}64_0"]
        np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=RTOL, atol=RTOL * np.abs(ref).max() * 1e-2, err_msg=n)
    # identical distributions: exactly zero
    assert (kl.detach().cpu().numpy()[0, 0, :4] == 0).all()


def test_kl_free_nats_fused_with_ties(B, O):
    """Fused KL + free nats + mask: values, sums and gradients incl. torch.maximum's 1/2 rule at the exact tie."""
    g = load_golden("kl_free_nats")
    Bn, Tz, Z = g["mu_q"].shape
    x_sl = torch.tensor([Tz, Tz - 4, 3])  # stride 1
    for i, fn in enumerate(g["free_nats"]):
        ins = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
        out = B.fused_elbo(None, None, x_sl, [B.KLLevel(*ins, stride=1)], beta=0.7, free_nats=float(fn), num_bins=2)
        out.loss.backward()
        ref = O.fused_elbo_value_and_grad(np.zeros((Bn, Tz)), np.zeros((Bn, Tz, 3)), x_sl.numpy(), [], 0.7, 1, 256)
        lv = dict(mu_q=g["mu_q"], sd_q=g["sd_q"], mu_p=g["mu_p"], sd_p=g["sd_p"], stride=1, free_nats=float(fn))
        m = O.sequence_mask(x_sl.numpy(), max_len=Tz)[..., None].astype(np.float64)
        kl, kl_fn, grads = O.kl_value_and_grad(*[g[n].astype(np.float64) for n in ("mu_q", "sd_q", "mu_p", "sd_p")],
                                               free_nats=float(fn), gout=m * (0.7 / float(x_sl.sum())))
        assert_sums_close(out.kl.cpu().numpy(), (kl * m).sum((1, 2)), "kl rows")
        assert_sums_close(out.kl_fn.cpu().numpy(), (kl_fn * m).sum((1, 2)), "kl_fn rows", rtol=2e-6)
        assert_sums_close(out.loss.item(), 0.7 * (kl_fn * m).sum() / float(x_sl.sum()), "loss", rtol=2e-6)
        tie = np.zeros(g["mu_q"].shape, bool)
        if i == 3:  # an exact tie of the reference's fp32 formula; in fp64 (the oracle here) it is not a tie
            tie[tuple(g["tie_index"])] = True
        for t, r in zip(ins, grads):
            np.testing.assert_allclose(t.grad.cpu().numpy()[~tie], r[~tie], rtol=RTOL, atol=RTOL * 1e-2 * np.abs(r).max())
    # the tie element: gradient is exactly half of the unconstrained one
    i, j, k = g["tie_index"]
    ins = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
    full = torch.tensor([Tz] * Bn)
    B.fused_elbo(None, None, full, [B.KLLevel(*ins, stride=1)], 1.0, float(g["free_nats"][3]), num_bins=2).loss.backward()
    ins0 = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
    B.fused_elbo(None, None, full, [B.KLLevel(*ins0, stride=1)], 1.0, 0.0, num_bins=2).loss.backward()
    # the CUDA KL value must hit the fp32 tie bit-exactly for this to hold; it is formulated differently from the
    # reference (log1p form), so only check the rule when it does
    kl_dev = B.kl_divergence_gaussian(*[t.detach() for t in ins])[i, j, k].item()
    if np.float32(kl_dev) == np.float32(g["free_nats"][3] / Z):
        assert math.isclose(ins[0].grad[i, j, k].item(), 0.5 * ins0[0].grad[i, j, k].item(), rel_tol=1e-6)


def _check_dmol_grads_fused(graw, g, K):
    x_sl = g["x_sl"]
    gabs = np.full(g["raw"].shape[:2], 1.0 / x_sl.sum()).reshape(-1)
    assert_grads_close(graw.reshape(-1, 3 * K), g["graw64"].reshape(-1, 3 * K), K, gabs, "d loss / d raw")


@pytest.mark.parametrize("name", ["elbo_vrnn_a", "elbo_vrnn_b", "elbo_srnn_a", "elbo_srnn_b"])
def test_vrnn_srnn_compute_elbo(name, B):
    g = load_golden(name)
    K, nb = int(g["K"]), int(g["num_bins"])
    fn = B.vrnn_compute_elbo if "vrnn" in name else B.srnn_compute_elbo
    self = SimpleNamespace(likelihood=B.DiscretizedLogisticMixtureDense(3 * K, 1, K, nb))
    raw = cu(g["raw"]).requires_grad_(True)
    ins = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
    kld = B.kl_divergence_gaussian(*ins)
    x_sl = torch.as_tensor(g["x_sl"])
    loss, elbo, logp, kl, mask = fn(self, cu(g["y"]).unsqueeze(-1), B.DMoLParams(raw, K, 1, -7.0), kld, x_sl,
                                    int(g["stride"]), float(g["beta"]), float(g["free_nats"]))
    loss.backward()
    for t in (loss, elbo, logp, kl, mask):
        assert t.dtype == torch.float64  # the reference's float64 quirk (vrnn.py:266)
    assert mask.shape == (len(x_sl), int(x_sl.max()))
    assert_sums_close(loss.item(), g["loss64"], "loss")
    assert_sums_close(elbo.detach().cpu().numpy(), g["elbo64"], "elbo")
    assert_sums_close(logp.detach().cpu().numpy(), g["logp64"], "log_prob")
    assert_sums_close(kl.detach().cpu().numpy(), g["kl64"], "kl", rtol=2e-6)
    _check_dmol_grads_fused(raw.grad.cpu().numpy(), g, K)
    for t, n in zip(ins, ("mu_q", "sd_q", "mu_p", "sd_p")):
        ref = g[f"g_{n}64"]
        np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=RTOL, atol=RTOL * 1e-2 * np.abs(ref).max(), err_msg=n)
    assert math.isclose(B.bits_per_dim(elbo, x_sl), -g["elbo64"].sum() / math.log(2) / g["x_sl"].sum(), rel_tol=1e-6)


def test_cwvae_compute_elbo(B):
    g = load_golden("elbo_cwvae")
    K, nb = int(g["K"]), int(g["num_bins"])
    ostr = [int(s) for s in g["overall_strides"]]
    self = SimpleNamespace(likelihood=B.DiscretizedLogisticMixtureDense(3 * K, 1, K, nb), num_levels=3, overall_strides=ostr)
    raw = cu(g["raw"]).requires_grad_(True)
    x_sl = torch.as_tensor(g["x_sl"])
    T = g["y"].shape[1]
    lv = [[cu(g[f"{n}_{l}"]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")] for l in range(3)]
    klds = [B.kl_divergence_gaussian(*ins) for ins in lv]
    seq_mask = B.sequence_mask(x_sl, max_len=T, device="cuda")
    level_masks = [B.sequence_mask((x_sl / s).ceil().int(), max_len=T // s, device="cuda") for s in ostr]
    loss, elbo, logp, kld, kld_l = B.cwvae_compute_elbo(self, cu(g["y"]).unsqueeze(-1), seq_mask, level_masks, x_sl,
                                                        B.DMoLParams(raw, K, 1, -7.0), klds, float(g["beta"]),
                                                        float(g["free_nats"]))
    loss.backward()
    assert elbo.dtype == torch.float32 and loss.dtype == torch.float32  # clockwork_vae uses bool masks
    np.testing.assert_allclose(loss.item(), g["loss64"], rtol=1e-6)
    np.testing.assert_allclose(elbo.detach().cpu().numpy(), g["elbo64"], rtol=1e-6)
    np.testing.assert_allclose(kld.detach().cpu().numpy(), g["kl64"], rtol=2e-6)
    for l in range(3):
        np.testing.assert_allclose(kld_l[l].detach().cpu().numpy(), g[f"kl_l{l}_64"], rtol=2e-6)
        for t, n in zip(lv[l], ("mu_q", "sd_q", "mu_p", "sd_p")):
            ref = g[f"g_{n}_{l}_64"]
            np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=RTOL, atol=RTOL * 1e-2 * np.abs(ref).max())
    _check_dmol_grads_fused(raw.grad.cpu().numpy(), g, K)


def test_stcn_compute_loss(B):
    g = load_golden("elbo_stcn")
    K, nb, n = int(g["K"]), int(g["num_bins"]), int(g["n_latents"])
    self = SimpleNamespace(likelihood_module=B.DiscretizedLogisticMixtureDense(3 * K, 1, K, nb),
                           n_stack_frames=int(g["n_stack_frames"]), top_down=True, n_latents=n)
    raw = cu(g["raw"]).requires_grad_(True)
    lv = [[cu(g[f"{nm}_{l}"]).requires_grad_(True) for nm in ("mu_q", "sd_q", "mu_p", "sd_p")] for l in range(n)]
    mu_q, sd_q, mu_p, sd_p = ([ins[i] for ins in lv] for i in range(4))
    loss, elbo, logp, kld, klds = B.stcn_compute_loss(self, cu(g["y"]).unsqueeze(-1), torch.as_tensor(g["x_sl"]),
                                                      B.DMoLParams(raw, K, 1, -7.0), mu_p, sd_p, mu_q, sd_q, None,
                                                      float(g["free_nats"]), float(g["beta"]))
    loss.backward()
    assert elbo.dtype == torch.float32
    np.testing.assert_allclose(loss.item(), g["loss64"], rtol=1e-6)
    np.testing.assert_allclose(elbo.detach().cpu().numpy(), g["elbo64"], rtol=1e-6)
    for l in range(n):
        np.testing.assert_allclose(klds[l].detach().cpu().numpy(), g[f"kl_l{l}_64"], rtol=2e-6)
        for t, nm in zip(lv[l], ("mu_q", "sd_q", "mu_p", "sd_p")):
            ref = g[f"g_{nm}_{l}_64"]
            np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=RTOL, atol=RTOL * 1e-2 * np.abs(ref).max())
    _check_dmol_grads_fused(raw.grad.cpu().numpy(), g, K)


def test_stcn_compute_loss_bottom_up(B):
    """top_down=False (stcn.py:288): Monte-Carlo KL through the Gaussian log-density kernels + the fused reduction."""
    g = load_golden("elbo_stcn_bottom_up")
    K, nb, n = int(g["K"]), int(g["num_bins"]), int(g["n_latents"])
    self = SimpleNamespace(likelihood_module=B.DiscretizedLogisticMixtureDense(3 * K, 1, K, nb),
                           n_stack_frames=int(g["n_stack_frames"]), top_down=False, n_latents=n)
    raw = cu(g["raw"]).requires_grad_(True)
    lv = [[cu(g[f"{nm}_{l}"]).requires_grad_(True) for nm in ("mu_q", "sd_q", "mu_p", "sd_p")] for l in range(n)]
    mu_q, sd_q, mu_p, sd_p = ([ins[i] for ins in lv] for i in range(4))
    z = [cu(g[f"z_{l}"]).requires_grad_(True) for l in range(n)]     # z = rsample(q) carries gradient in the model
    loss, elbo, logp, kld, klds = B.stcn_compute_loss(self, cu(g["y"]).unsqueeze(-1), torch.as_tensor(g["x_sl"]),
                                                      B.DMoLParams(raw, K, 1, -7.0), mu_p, sd_p, mu_q, sd_q, z,
                                                      float(g["free_nats"]), float(g["beta"]))
    loss.backward()
    assert elbo.dtype == torch.float32
    np.testing.assert_allclose(loss.item(), g["loss64"], rtol=2e-6)
    np.testing.assert_allclose(elbo.detach().cpu().numpy(), g["elbo64"], rtol=2e-6)
    np.testing.assert_allclose(kld.detach().cpu().numpy(), g["kl64"], rtol=1e-5, atol=1e-5)
    for l in range(n):
        ref_kl = g[f"kl_l{l}_64"]                        # a sum of signed MC terms: absolute floor from its own scale
        np.testing.assert_allclose(klds[l].detach().cpu().numpy(), ref_kl, rtol=1e-5, atol=1e-6 * np.abs(ref_kl).max() + 1e-5)
        for t, nm in zip(lv[l] + [z[l]], ("mu_q", "sd_q", "mu_p", "sd_p", "z")):
            ref = g[f"g_{nm}_{l}_64"]
            np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=RTOL, atol=RTOL * 1e-2 * np.abs(ref).max())
    _check_dmol_grads_fused(raw.grad.cpu().numpy(), g, K)


def test_wavenet_compute_loss(B):
    g = load_golden("elbo_wavenet")
    K, nb = int(g["K"]), int(g["num_bins"])
    self = SimpleNamespace(likelihood=B.DiscretizedLogisticMixtureDense(3 * K, 1, K, nb))
    raw = cu(g["raw"]).requires_grad_(True)
    x_sl = torch.as_tensor(g["x_sl"])
    loss, logp, twise = B.wavenet_compute_loss(self, cu(g["y"]).unsqueeze(-1), x_sl, B.DMoLParams(raw, K, 1, -7.0))
    loss.backward()
    assert logp.dtype == torch.float32 and twise.shape == g["logp_twise64"].shape
    np.testing.assert_allclose(loss.item(), g["loss64"], rtol=1e-6)
    np.testing.assert_allclose(logp.detach().cpu().numpy(), g["logp64"], rtol=1e-6)
    assert_values_close(twise.cpu().numpy(), g["logp_twise64"], "log_prob_twise")
    assert (twise.cpu().numpy()[3, 1:] == 0).all()  # x_sl = 1: everything after the first sample is masked
    _check_dmol_grads_fused(raw.grad.cpu().numpy(), g, K)
    np.testing.assert_allclose(B.bits_per_dim(logp, x_sl), float(g["bpd32"]), rtol=1e-6)


@pytest.mark.parametrize("bits", [8, 16])
def test_quantize_bit_exact(bits, B):
    g = load_golden("quantize")
    q = B.Quantize(bits=bits).cuda()
    assert np.array_equal(q.boundaries.cpu().numpy(), g[f"boundaries_{bits}"])
    pcm = torch.arange(-32768, 32768, dtype=torch.float32) / 32768
    assert np.array_equal(q(pcm.cuda()).cpu().numpy(), g[f"pcm_idx_{bits}"])
    assert np.array_equal(q(cu(g["rnd"])).cpu().numpy(), g[f"rnd_idx_{bits}"])
    assert np.array_equal(q(cu(g[f"boundaries_{bits}"])).cpu().numpy(), g[f"bnd_idx_{bits}"])
    grid = torch.arange(0, 2 ** bits, dtype=torch.float32) / (2 ** bits - 1) * 2 - 1
    out = q(grid.cuda())
    assert out.dtype == torch.int64 and np.array_equal(out.cpu().numpy(), g[f"grid_idx_{bits}"])


# ---- edge cases: ragged / empty / unaligned shapes ------------------------------------------------------------------
@pytest.mark.parametrize("T,K", [(1, 10), (7, 10), (127, 10), (129, 10), (255, 3), (301, 1), (130, 2), (64, 30)])
def test_ragged_and_unaligned_shapes(T, K, B, O):
    """Odd T makes row slabs 8- or 4-byte aligned only (the non-TMA path), T < tile and T = tile+1 hit the tail tile;
    lengths include 0 and T."""
    rng = np.random.default_rng(T * 31 + K)
    Bn, nb = 5, 65536
    x_sl = np.array([T, max(T - 1, 0), T // 2, 1 if T > 1 else 0, 0])
    if x_sl.sum() == 0:
        x_sl[0] = T
    y = (rng.integers(0, nb, (Bn, T)) / (nb - 1) * 2 - 1).astype(np.float32)
    raw = rng.normal(size=(Bn, T, 3 * K)).astype(np.float32)
    raw[..., K:2 * K] = y[..., None] + 0.05 * rng.normal(size=(Bn, T, K))
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    for skip in (False, True):
        r = cu(raw).requires_grad_(True)
        out = B.fused_elbo(cu(y), B.DMoLParams(r, K, 1, -7.0), torch.as_tensor(x_sl), (), 1.0, 0.0, num_bins=nb,
                           want_twise=True, skip_padded=skip)
        out.loss.backward()
        ref = O.fused_elbo_value_and_grad(y, raw, x_sl, [], 1.0, K, nb)
        assert_sums_close(out.loss.item(), ref["loss"], "loss")
        assert_sums_close(out.log_prob.cpu().numpy(), ref["logp"], "rows")
        assert_values_close(out.log_prob_twise.cpu().numpy(), ref["lp_twise"], "twise")
        gabs = np.full(Bn * T, 1.0 / x_sl.sum())
        assert_grads_close(r.grad.cpu().numpy().reshape(-1, 3 * K), ref["graw"].reshape(-1, 3 * K), K, gabs)
        m = O.sequence_mask(x_sl, max_len=T)
        assert (r.grad.cpu().numpy()[~m] == 0).all()  # padded positions get exact zeros


def test_offset_view_inputs(B, O):
    """Parameters that are a non-16-byte-aligned view of a larger buffer (odd float offset)."""
    rng = np.random.default_rng(5)
    Bn, T, K, nb = 2, 200, 10, 65536
    big = torch.zeros(Bn * T * 3 * K + 3, device="cuda")
    raw_np = rng.normal(size=(Bn, T, 3 * K)).astype(np.float32)
    raw_np[..., 2 * K:] = raw_np[..., 2 * K:] * 2 - 4
    view = big[3:].view(Bn, T, 3 * K)
    view.copy_(cu(raw_np))
    assert view.data_ptr() % 16 != 0
    y = (rng.integers(0, nb, (Bn, T)) / (nb - 1) * 2 - 1).astype(np.float32)
    lik = B.DiscretizedLogisticMixtureDense(3, 1, K, nb)
    lp = lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(view, K, 1, -7.0))
    ref, _ = O.dmol_value_and_grad(y.reshape(-1).astype(np.float64), raw_np.reshape(-1, 3 * K).astype(np.float64), K, 1, nb)
    assert_values_close(lp.cpu().numpy().reshape(-1), ref, "offset view")


def test_empty_batch(B):
    lik = B.DiscretizedLogisticMixtureDense(3, 1, 10, 65536)
    lp = lik.log_prob(torch.zeros(0, 5, 1, device="cuda"), B.DMoLParams(torch.zeros(0, 5, 30, device="cuda"), 10, 1, -7.0))
    assert lp.shape == (0, 5)
    kl = B.kl_divergence_gaussian(*[torch.ones(0, 3, 4, device="cuda")] * 4)
    assert kl.shape == (0, 3, 4)


def test_cpu_tensors_raise(B):
    lik = B.DiscretizedLogisticMixtureDense(3, 1, 10, 65536)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lik.log_prob(torch.zeros(2, 5, 1), B.DMoLParams(torch.zeros(2, 5, 30), 10, 1, -7.0))


def test_out_of_range_targets_are_flagged(B):
    lik = B.DiscretizedLogisticMixtureDense(3, 1, 10, 65536)
    y = torch.zeros(2, 5, 1, device="cuda")
    y[1, 2, 0] = 1.5
    lik.log_prob(y, B.DMoLParams(torch.zeros(2, 5, 30, device="cuda"), 10, 1, -7.0))
    with pytest.raises(AssertionError):
        B.check_input_range()
    B.check_input_range()  # flag was reset


def test_nan_propagates(B):
    """NaN parameters poison that sample (and, through `* mask`, its utterance sum) like in the reference."""
    raw = torch.zeros(2, 130, 30, device="cuda")
    raw[0, 3, 12] = float("nan")
    out = B.fused_elbo(torch.zeros(2, 130, device="cuda"), B.DMoLParams(raw, 10, 1, -7.0), torch.tensor([130, 130]), (),
                       num_bins=65536, want_twise=True)
    tw = out.log_prob_twise.cpu().numpy()
    assert np.isnan(tw[0, 3]) and np.isfinite(np.delete(tw.reshape(-1), 3)).all()
    assert np.isnan(out.log_prob[0].item()) and np.isfinite(out.log_prob[1].item())


def test_upstream_grad_scale_and_double_backward(B):
    """loss * c backpropagates c * grads through the early-exit scale kernel (AMP GradScaler path)."""
    g = load_golden("elbo_srnn_a")
    K, nb = 10, 65536

    def run(scale):
        raw = cu(g["raw"]).requires_grad_(True)
        ins = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
        out = B.fused_elbo(cu(g["y"]), B.DMoLParams(raw, K, 1, -7.0), torch.as_tensor(g["x_sl"]),
                           [B.KLLevel(*ins, stride=int(g["stride"]))], 0.5, 0.0625, num_bins=nb)
        (out.loss * scale).backward()
        return raw.grad, ins[1].grad, out

    g1, k1, out = run(1.0)
    g3, k3, _ = run(1024.0)
    torch.testing.assert_close(g3, g1 * 1024.0, rtol=1e-6, atol=0)
    torch.testing.assert_close(k3, k1 * 1024.0, rtol=1e-6, atol=0)
    with pytest.raises(RuntimeError, match="backward called twice"):
        out.loss.backward()


def test_no_grad_eval_path(B):
    g = load_golden("elbo_vrnn_b")
    raw = cu(g["raw"]).requires_grad_(True)
    with torch.no_grad():
        out = B.fused_elbo(cu(g["y"]), B.DMoLParams(raw, 10, 1, -7.0), torch.as_tensor(g["x_sl"]), (), num_bins=65536)
    assert not out.loss.requires_grad
    assert_sums_close(out.log_prob.cpu().numpy(), g["logp64"], "eval rows")
    m = B.elbo_metrics(out)
    assert math.isclose(m.bpd, -g["logp64"].sum() / math.log(2) / g["x_sl"].sum(), rel_tol=1e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("K", [10, 5, 30])
def test_half_precision_parameters(dtype, K, B, O):
    """Under AMP the Linear output arrives in bf16/fp16 and the kernels read it as it is (arithmetic stays fp32): values
    equal the fp32 call on the same rounded parameters to the parity tolerance, gradients come back in the parameter dtype and
    match the oracle on the rounded parameters to the dtype's resolution."""
    rng = np.random.default_rng(K)
    Bn, T, nb = 3, 300, 65536
    y = (rng.integers(0, nb, (Bn, T)) / (nb - 1) * 2 - 1).astype(np.float32)
    raw = rng.normal(size=(Bn, T, 3 * K)).astype(np.float32)
    raw[..., K:2 * K] = y[..., None] + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    raw_h = cu(raw).to(dtype)
    x_sl = torch.tensor([300, 211, 17])
    lik = B.DiscretizedLogisticMixtureDense(3, 1, K, nb)
    with torch.autocast("cuda", dtype=dtype):
        lp_h = lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(raw_h, K, 1, -7.0))
    lp_f = lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(raw_h.float(), K, 1, -7.0))
    # same rounded parameters: the 16-bit kernels evaluate the mixture in the linear domain (blvm_math.cuh: dmol_sample_lin), the fp32
    # kernel in the log domain; both are within the tolerance of the fp64 anchor, hence within 2x of it of each other
    assert lp_h.dtype == torch.float32
    assert_values_close(lp_h.cpu().numpy(), lp_f.cpu().numpy().astype(np.float64), "16-bit parameters vs fp32 call on the same values", rtol=2e-5)
    # fused op: loss * scale backward (GradScaler-style), gradient dtype == parameter dtype
    scale = 4096.0
    r = raw_h.clone().requires_grad_(True)
    out = B.fused_elbo(cu(y), B.DMoLParams(r, K, 1, -7.0), x_sl, (), num_bins=nb)
    (out.loss * scale).backward()
    assert r.grad.dtype == dtype
    ref = O.fused_elbo_value_and_grad(y, raw_h.float().cpu().numpy(), x_sl.numpy(), [], 1.0, K, nb)
    assert_sums_close(out.loss.item(), ref["loss"], "loss on rounded parameters")
    g = r.grad.float().cpu().numpy().astype(np.float64) / scale
    gref = ref["graw"]
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    tol = eps * np.abs(gref) + 1e-4 * np.abs(gref).max(-1, keepdims=True) + (6e-8 / scale if dtype == torch.float16 else 0)
    assert (np.abs(g - gref) <= tol).all()
    # generic autograd path (recompute backward) in the same dtype
    r2 = raw_h.clone().requires_grad_(True)
    lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(r2, K, 1, -7.0)).sum().backward()
    assert r2.grad.dtype == dtype and torch.isfinite(r2.grad.float()).all()
    # a K without a register kernel is upcast transparently
    r7 = torch.randn(2, 50, 21, device="cuda").to(dtype).requires_grad_(True)
    lp7 = B.DiscretizedLogisticMixtureDense(3, 1, 7, nb).log_prob(torch.zeros(2, 50, 1, device="cuda"), B.DMoLParams(r7, 7, 1, -7.0))
    lp7.sum().backward()
    assert r7.grad.dtype == dtype


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("K", [8, 12, 16, 20])
def test_wide_rows_and_rotated_section_walk(K, dtype, B, O):
    """Rows that are a whole number of 16 / 8-byte vectors are lifted with 128 / 64-bit shared-memory accesses, and where a row
    is an EVEN number of 16-byte vectors (fp32 K = 8 / 16, 16-bit K = 16) each lane walks the three K-element sections in a
    lane-dependent rotated order (dmol_kernels.cuh: RowRot; removes a 2 / 4-way bank conflict).  The permutation is the same
    in all sections and the likelihood is symmetric in the component index: values and gradients against the oracle, every
    gradient lands on its own component, ragged tails and a second tile included."""
    rng = np.random.default_rng(100 + K)
    Bn, T, nb = 3, 517, 65536
    y = (rng.integers(0, nb, (Bn, T)) / (nb - 1) * 2 - 1).astype(np.float32)
    raw = rng.normal(size=(Bn, T, 3 * K)).astype(np.float32)
    raw[..., K:2 * K] = y[..., None] + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    raw[..., :K] += np.arange(K, dtype=np.float32) * 0.37            # components are distinguishable
    raw_h = cu(raw).to(dtype)
    x_sl = torch.tensor([T, 300, 17])
    r = raw_h.clone().requires_grad_(True)
    scale = 4096.0 if dtype != torch.float32 else 1.0
    out = B.fused_elbo(cu(y), B.DMoLParams(r, K, 1, -7.0), x_sl, (), num_bins=nb, want_twise=True)
    (out.loss * scale).backward()
    ref = O.fused_elbo_value_and_grad(y, raw_h.float().cpu().numpy(), x_sl.numpy(), [], 1.0, K, nb)
    assert_sums_close(out.loss.item(), ref["loss"], "loss")
    g = r.grad.float().cpu().numpy().astype(np.float64) / scale
    gref = ref["graw"]
    if dtype == torch.float32:
        n = float(x_sl.sum())
        mask = O.sequence_mask(x_sl.numpy(), max_len=T).reshape(-1)
        assert_grads_close(g, gref, K, mask / n, "d/d raw")
    else:
        eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
        tol = eps * np.abs(gref) + 1e-4 * np.abs(gref).max(-1, keepdims=True) + (6e-8 / scale if dtype == torch.float16 else 0)
        assert (np.abs(g - gref) <= tol).all()
    # the same rows through the forward-only kernel and the 16-bit rows against the fp32 kernel on the same rounded values
    lik = B.DiscretizedLogisticMixtureDense(3, 1, K, nb)
    lp_h = lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(raw_h, K, 1, -7.0))
    lp_f = lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(raw_h.float(), K, 1, -7.0))
    assert_values_close(lp_h.cpu().numpy(), lp_f.cpu().numpy().astype(np.float64), "16-bit rows vs fp32 rows", rtol=2e-6)
    tw = out.log_prob_twise.cpu().numpy() if hasattr(out, "log_prob_twise") else None
    if tw is not None:
        m = O.sequence_mask(x_sl.numpy(), max_len=T)
        assert_values_close(tw[m], lp_h.cpu().numpy()[m], "fused step vs forward-only kernel", rtol=2e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("K", [2, 5, 10])
def test_linear_domain_fallback(K, dtype, B, O):
    """The 16-bit kernels evaluate the mixture in the linear domain (dmol_sample_lin) and fall back to the log-domain body for
    samples the product cannot represent: every component tens to thousands of scales away (all exp(-|m|) underflow), NaN /
    inf parameters, the two edge bins.  Half of the samples here are such cases; values and gradients against the oracle on the
    rounded parameters, and non-finite rows exactly like the fp32 (log-domain only) kernel."""
    rng = np.random.default_rng(7 + K)
    Bn, T, nb = 2, 1100, 65536
    y = (rng.integers(0, nb, (Bn, T)) / (nb - 1) * 2 - 1).astype(np.float32)
    y[0, :8] = np.array([-1.0, 1.0, 2 / nb - 1, 1 - 2 / nb, -1.0, 1.0, 0.0, 0.5], np.float32)     # edge bins and their neighbours
    raw = rng.normal(size=(Bn, T, 3 * K)).astype(np.float32)
    raw[..., K:2 * K] = y[..., None] + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    far = rng.random((Bn, T)) < 0.5                                   # every component far away with a tiny scale
    shift = rng.choice([0.05, 0.3, 1.5], size=(Bn, T, K)) * rng.choice([-1.0, 1.0], size=(Bn, T, K))
    raw[..., K:2 * K] = np.where(far[..., None], y[..., None] + shift, raw[..., K:2 * K])
    raw[..., 2 * K:] = np.where(far[..., None], rng.uniform(-9.0, -5.5, (Bn, T, K)), raw[..., 2 * K:])
    raw[..., :K] += np.where(far[..., None], rng.normal(0, 20, (Bn, T, K)), 0.0)      # widely spread logits as well
    raw_h = cu(raw).to(dtype)
    x_sl = torch.tensor([T, 700])
    scale = 1024.0
    r = raw_h.clone().requires_grad_(True)
    out = B.fused_elbo(cu(y), B.DMoLParams(r, K, 1, -7.0), x_sl, (), num_bins=nb, want_twise=True)
    (out.loss * scale).backward()
    ref = O.fused_elbo_value_and_grad(y, raw_h.float().cpu().numpy(), x_sl.numpy(), [], 1.0, K, nb)
    m = O.sequence_mask(x_sl.numpy(), max_len=T)
    tw = out.log_prob_twise.cpu().numpy().astype(np.float64)
    assert_values_close(tw[m], ref["lp_twise"][m], "per-sample log-prob, far-out samples included")
    assert float(np.abs(ref["lp_twise"][m & far]).max()) > 100.0        # the fallback regime really is exercised
    assert_sums_close(out.loss.item(), ref["loss"], "loss")
    g = r.grad.float().cpu().numpy().astype(np.float64) / scale
    gref = ref["graw"]
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    # absolute floor: d/d logit = g (resp - softmax) cancels to ~1e-7 g when one logit dominates (parity.py uses the same 1e-6 g)
    tol = eps * np.abs(gref) + 1e-4 * np.abs(gref).max(-1, keepdims=True) + (6e-8 / scale if dtype == torch.float16 else 0) + 1e-6 / float(x_sl.sum())
    ratio = np.abs(g - gref) / tol
    w = np.unravel_index(np.argmax(ratio), ratio.shape)
    assert ratio.max() <= 1.0, (f"worst err/tol {ratio.max():.3g} at {w}: ours {g[w]:.6g} ref {gref[w]:.6g}, far={far[w[0], w[1]]}, y={y[w[0], w[1]]}, "
                                f"row={raw_h[w[0], w[1]].float().cpu().numpy()}, ref row grad={gref[w[0], w[1]]}, ours={g[w[0], w[1]]}")
    # non-finite parameters propagate exactly like in the fp32 kernel (log domain only)
    bad = raw_h.clone()
    bad[0, 10, 0] = float("nan")            # NaN logit
    bad[0, 11, K] = float("nan")            # NaN location
    bad[0, 12, 2 * K] = float("nan")        # NaN log-scale
    bad[0, 13, 0] = float("inf")            # +inf logit
    bad[0, 14, 1] = -float("inf")           # -inf logit: that component simply has zero weight
    lik = B.DiscretizedLogisticMixtureDense(3, 1, K, nb)
    lp_h = lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(bad, K, 1, -7.0))
    lp_f = lik.log_prob(cu(y).unsqueeze(-1), B.DMoLParams(bad.float(), K, 1, -7.0))
    fin_h, fin_f = torch.isfinite(lp_h), torch.isfinite(lp_f)
    assert torch.equal(fin_h, fin_f) and torch.equal(torch.isnan(lp_h), torch.isnan(lp_f))
    assert_values_close(lp_h[fin_h].cpu().numpy(), lp_f[fin_f].cpu().numpy().astype(np.float64), "finite rows", rtol=2e-5)
    assert bool(torch.isfinite(lp_h[0, 14]))


def _random_cases():
    rng = np.random.default_rng(int(os.environ.get("BLVM_TEST_SEED", "20260")))   # another seed = another set of shapes
    cases = []
    for K in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 30):
        for dt in ("float32", "bfloat16", "float16"):
            Bn = int(rng.integers(1, 5))
            T = int(rng.choice([int(rng.integers(1, 200)), int(rng.integers(200, 2600)), 4 * int(rng.integers(50, 700))]))
            cases.append((K, dt, Bn, T, int(rng.integers(0, 2 ** 31))))
    return cases


@pytest.mark.parametrize("K,dt,Bn,T,seed", _random_cases())
def test_random_shapes_against_oracle(K, dt, Bn, T, seed, B, O):
    """Every instantiated mixture size x parameter dtype at a random (batch, length) with ragged lengths, one KL level of a random
    stride: the whole fused step (loss, row sums, per-sample log-prob, every gradient) against the oracle on the rounded parameters.
    Covers the kernel variants no dedicated test pins to a shape: tile / stream kernel by alignment, tails shorter than a tile,
    rotated and packed row walks, linear-domain evaluation with its fallback, 64/128-bit row accesses."""
    dtype = getattr(torch, dt)
    rng = np.random.default_rng(seed)
    nb = 65536
    y = (rng.integers(0, nb, (Bn, T)) / (nb - 1) * 2 - 1).astype(np.float32)
    raw = rng.normal(size=(Bn, T, 3 * K)).astype(np.float32)
    raw[..., K:2 * K] = y[..., None] + rng.choice([0.02, 0.1, 0.5], size=(Bn, T, 1)) * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    raw_h = cu(raw).to(dtype)
    x_sl = torch.from_numpy(rng.integers(1, T + 1, Bn))
    x_sl[int(rng.integers(0, Bn))] = T
    S = int(rng.choice([1, 3, 16, 64]))
    Tz, Z = -(-T // S), int(rng.choice([1, 3, 8]))
    klin = [rng.normal(size=(Bn, Tz, Z)).astype(np.float32), (rng.random((Bn, Tz, Z)) + 0.1).astype(np.float32),
            rng.normal(size=(Bn, Tz, Z)).astype(np.float32), (rng.random((Bn, Tz, Z)) + 0.1).astype(np.float32)]
    kl_t = [cu(t).requires_grad_(True) for t in klin]
    r = raw_h.clone().requires_grad_(True)
    scale = 1.0 if dtype == torch.float32 else 1024.0
    out = B.fused_elbo(cu(y), B.DMoLParams(r, K, 1, -7.0), x_sl, [B.KLLevel(*kl_t, stride=S)], 0.7, 0.5, num_bins=nb, want_twise=True)
    (out.loss * scale).backward()
    B.check_input_range()
    ref = O.fused_elbo_value_and_grad(y, raw_h.float().cpu().numpy(), x_sl.numpy(),
                                      [dict(mu_q=klin[0], sd_q=klin[1], mu_p=klin[2], sd_p=klin[3], stride=S, free_nats=0.5)], 0.7, K, nb)
    m = O.sequence_mask(x_sl.numpy(), max_len=T)
    assert_sums_close(out.loss.item(), ref["loss"], "loss")
    assert_sums_close(out.log_prob.cpu().numpy(), ref["logp"], "row log-prob")
    assert_sums_close(out.kl.cpu().numpy(), ref["kl"], "row KL", rtol=2e-6)
    tw = out.log_prob_twise.cpu().numpy().astype(np.float64)
    assert_values_close(tw[m], ref["lp_twise"][m], "per-sample log-prob")
    assert (tw[~m] == 0).all()
    g = r.grad.float().cpu().numpy().astype(np.float64) / scale
    gref = ref["graw"]
    n = float(x_sl.sum())
    if dtype == torch.float32:
        assert_grads_close(g, gref, K, m.reshape(-1) / n, "d/d raw")
    else:
        eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
        tol = eps * np.abs(gref) + 1e-4 * np.abs(gref).max(-1, keepdims=True) + (6e-8 / scale if dtype == torch.float16 else 0) + 1e-6 / n
        ratio = np.abs(g - gref) / tol
        assert ratio.max() <= 1.0, f"worst err/tol {ratio.max():.3g} at {np.unravel_index(np.argmax(ratio), ratio.shape)}"
    assert (g[~m] == 0).all()
    for t, gr_ref, nm in zip(kl_t, ref["gkl"][0], ("mu_q", "sd_q", "mu_p", "sd_p")):
        np.testing.assert_allclose(t.grad.cpu().numpy() / scale, gr_ref, rtol=2e-5, atol=1e-7 * np.abs(gr_ref).max() + 1e-12, err_msg=nm)


def test_fp16_gradients_do_not_underflow_with_loss_scale(B, O):
    """fp16 parameters: d loss/d raw ~ 1/sum(x_sl) ~ 1e-7 is below fp16's subnormal range; with the GradScaler's loss
    scale applied inside the backward launch (device scalar) the scaled gradients are representable."""
    g = load_golden("elbo_wavenet")
    K, nb = 10, 65536
    x_sl = torch.full((4,), 80)
    raw = cu(np.tile(g["raw"], (1, 1, 1))).to(torch.float16).requires_grad_(True)
    out = B.fused_elbo(cu(g["y"]), B.DMoLParams(raw, K, 1, -7.0), x_sl, (), num_bins=nb, denom=4.0e6)
    scale = torch.tensor(65536.0, device="cuda", dtype=torch.float64)
    (out.loss * scale).backward()
    gr = raw.grad.float()
    ref = O.fused_elbo_value_and_grad(g["y"], raw.detach().float().cpu().numpy(), x_sl.numpy(), [], 1.0, K, nb)
    gref = ref["graw"] * (float(x_sl.sum()) / 4.0e6) * 65536.0
    big = np.abs(gref) > 1e-3
    assert big.sum() > 1000
    np.testing.assert_allclose(gr.cpu().numpy()[big], gref[big], rtol=2e-3)
    assert (gr != 0).float().mean().item() > 0.5


# ---- full benchmark size: size-independent properties ---------------------------------------------------------------
def test_full_size_properties(B, O):
    """BASELINE config 5 (B=256, T=16000, K=10): determinism, twise/row-sum consistency, component-permutation
    invariance, gradient identities, and a row subset against the oracle."""
    torch.manual_seed(1234)
    Bn, T, K, nb, S, Z = 256, 16000, 10, 65536, 64, 64
    dev = "cuda"
    y = (torch.randint(0, nb, (Bn, T), device=dev).float() / (nb - 1) * 2 - 1)
    raw = torch.randn(Bn, T, 3 * K, device=dev)
    raw[..., K:2 * K] = y.unsqueeze(-1) + 0.1 * torch.randn(Bn, T, K, device=dev)
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    x_sl = (T * (0.5 + 0.5 * torch.rand(Bn))).long()
    x_sl[0] = T
    Tz = T // S
    kl = [torch.randn(Bn, Tz, Z, device=dev), torch.nn.functional.softplus(torch.randn(Bn, Tz, Z, device=dev)) + 1e-3,
          torch.randn(Bn, Tz, Z, device=dev), torch.nn.functional.softplus(torch.randn(Bn, Tz, Z, device=dev)) + 1e-3]

    def run(raw_in, kl_in):
        r = raw_in.clone().requires_grad_(True)
        k = [t.clone().requires_grad_(True) for t in kl_in]
        out = B.fused_elbo(y, B.DMoLParams(r, K, 1, -7.0), x_sl, [B.KLLevel(*k, stride=S)], 0.5, 0.0625, num_bins=nb,
                           want_twise=True)
        out.loss.backward()
        return out, r.grad, [t.grad for t in k]

    out1, g1, kg1 = run(raw, kl)
    out2, g2, kg2 = run(raw, kl)
    # (1) deterministic: bit-identical run to run (no atomics)
    assert torch.equal(out1.elbo, out2.elbo) and torch.equal(g1, g2) and torch.equal(out1.loss, out2.loss)
    # (2) per-sample values are consistent with the per-utterance sums, and masked positions are zero
    tw = out1.log_prob_twise.double()
    torch.testing.assert_close(tw.sum(1), out1.log_prob, rtol=1e-9, atol=0)
    mask = torch.arange(T, device=dev)[None] < x_sl.to(dev)[:, None]
    assert (tw[~mask] == 0).all() and (g1[~mask] == 0).all()
    # (3) loss identity and bpd
    loss = -(out1.log_prob - 0.5 * out1.kl_fn).sum() / x_sl.sum()
    torch.testing.assert_close(out1.loss, loss.to(out1.loss.device), rtol=1e-12, atol=0)
    assert (out1.kl_fn >= out1.kl - 1e-9).all()
    # (4) softmax identity: d/d logits sums to zero per sample
    assert g1[..., :K].sum(-1).abs().max().item() < 1e-6 / x_sl.sum().item() * 50
    # (5) permuting the mixture components permutes the gradients and leaves every value unchanged (to rounding)
    perm = torch.randperm(K, device=dev)
    idx = torch.cat([perm, K + perm, 2 * K + perm])
    out3, g3, _ = run(raw[..., idx].contiguous(), kl)
    torch.testing.assert_close(out3.log_prob_twise, out1.log_prob_twise, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out3.log_prob, out1.log_prob, rtol=1e-7, atol=0)
    torch.testing.assert_close(g3, g1[..., idx], rtol=1e-4, atol=1e-6 / x_sl.sum().item())
    # (6) a row subset against the oracle (fp64), full length
    rows = [0, 17, 255]
    ref = O.fused_elbo_value_and_grad(y[rows].cpu().numpy(), raw[rows].cpu().numpy(), x_sl[rows].numpy(),
                                      [dict(mu_q=kl[0][rows].cpu().numpy(), sd_q=kl[1][rows].cpu().numpy(),
                                            mu_p=kl[2][rows].cpu().numpy(), sd_p=kl[3][rows].cpu().numpy(), stride=S,
                                            free_nats=0.0625)], 0.5, K, nb)
    assert_sums_close(out1.log_prob[rows].cpu().numpy(), ref["logp"], "rows vs oracle")
    assert_sums_close(out1.kl[rows].cpu().numpy(), ref["kl"], "kl vs oracle", rtol=2e-6)
    assert_values_close(out1.log_prob_twise[rows].cpu().numpy(), ref["lp_twise"], "twise vs oracle")
    scale = float(x_sl[rows].sum()) / float(x_sl.sum())  # oracle normalised by the subset's length
    gabs = np.full(len(rows) * T, 1.0 / float(x_sl.sum()))
    assert_grads_close(g1[rows].cpu().numpy().reshape(-1, 3 * K), ref["graw"].reshape(-1, 3 * K) * scale, K, gabs)


def test_config5_top_end_64bit_indexing(B, O):
    """BASELINE config 5's largest point: 256 x 128000 samples, K = 30 -> 2.95 G parameter elements (> 2^31), 11.8 GB of
    parameters + 11.8 GB of gradients.  Checks the last utterance (highest addresses) against the oracle, the sums
    against the per-sample values, and exact zeros in the padding."""
    Bn, T, K, nb = 256, 128000, 30, 65536
    dev = "cuda"
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~30 GB of free HBM")
    gen = torch.Generator(device=dev).manual_seed(7)
    y = torch.randint(0, nb, (Bn, T), device=dev, generator=gen).float() / (nb - 1) * 2 - 1
    raw = torch.empty(Bn, T, 3 * K, device=dev)
    for b0 in range(0, Bn, 32):  # fill in slices: keeps the temporary small
        sl = slice(b0, b0 + 32)
        raw[sl].normal_(generator=gen)
        raw[sl, :, K:2 * K] = y[sl].unsqueeze(-1) + 0.1 * raw[sl, :, K:2 * K]
        raw[sl, :, 2 * K:] = raw[sl, :, 2 * K:] * 2 - 4
    assert raw.numel() > 2 ** 31
    x_sl = torch.full((Bn,), T)
    x_sl[-1] = T - 12345
    raw.requires_grad_(True)
    out = B.fused_elbo(y, B.DMoLParams(raw, K, 1, -7.0), x_sl, (), num_bins=nb, want_twise=True)
    out.loss.backward()
    torch.testing.assert_close(out.log_prob_twise.double().sum(1), out.log_prob, rtol=1e-9, atol=0)
    last = Bn - 1
    sub = slice(T - 20000, T)   # the tail of the last utterance: highest addresses, straddles the mask edge
    ref_lp, ref_g = O.dmol_value_and_grad(y[last, sub].cpu().numpy().astype(np.float64),
                                          raw[last, sub].detach().cpu().numpy().astype(np.float64), K, 1, nb)
    m = (np.arange(T)[sub] < int(x_sl[last])).astype(np.float64)
    assert_values_close(out.log_prob_twise[last, sub].cpu().numpy(), ref_lp * m, "twise at the top of the address range")
    gs = -m / float(x_sl.sum())
    g = raw.grad[last, sub].cpu().numpy()
    assert_grads_close(g, ref_g * gs[:, None], K, np.abs(gs) + 1e-30, "grads at the top of the address range")
    assert (g[m == 0] == 0).all() and (out.log_prob_twise[last, sub].cpu().numpy()[m == 0] == 0).all()
    del raw, out
    torch.cuda.empty_cache()


def test_fused_sample_and_mode(B):
    """sample() + mode() from one read of the parameters.  mode is exact (argmax + gather, with the
    gather's gradient); the sample stream is Philox-based, so parity with the reference's torch RNG is distributional:
    mixture frequencies follow softmax(logits), each component is a logistic(loc, exp(log_scale)), values are clamped to
    [-1, 1] (blvm/utils/variational.py:282-349)."""
    from scipy import stats
    K, nb = 4, 65536
    lik = B.DiscretizedLogisticMixtureDense(3, 1, K, nb)
    # --- mode: exact, differentiable like the reference's gather
    raw = torch.randn(6, 50, 3 * K, device="cuda", requires_grad=True)
    p = B.DMoLParams(raw, K, 1, -7.0)
    n0 = B.launch_count()
    smp = lik.sample(p)
    mode = lik.mode(p)
    assert B.launch_count() - n0 == 1                      # one kernel served both calls
    assert smp.shape == (6, 50, 1) and mode.shape == (6, 50, 1) and not smp.requires_grad
    idx = raw[..., :K].argmax(-1, keepdim=True)
    ref_mode = torch.gather(raw[..., K:2 * K], -1, idx)
    assert torch.equal(mode.detach(), ref_mode.detach())
    w = torch.randn_like(mode)
    (g_ours,) = torch.autograd.grad((mode * w).sum(), raw)
    (g_ref,) = torch.autograd.grad((ref_mode * w).sum(), raw)
    assert torch.equal(g_ours, g_ref)
    assert float(smp.abs().max()) <= 1.0
    # --- sample statistics on one parameter row replicated N times
    N = 400000
    logits = torch.tensor([0.5, -1.0, 2.0, 0.0])
    locs = torch.tensor([-0.6, -0.2, 0.2, 0.6])
    ls = torch.tensor([-5.0, -4.5, -5.5, -4.0])
    row = torch.cat([logits, locs, ls]).cuda()
    torch.manual_seed(123)
    x = lik.sample(B.DMoLParams(row.expand(N, 3 * K).contiguous(), K, 1, -7.0)).squeeze(-1).double().cpu()
    comp = (x.unsqueeze(-1) - locs.double()).abs().argmin(-1)       # components are ~20 scales apart
    freq = torch.bincount(comp, minlength=K).double() / N
    pi = torch.softmax(logits.double(), -1)
    assert ((freq - pi).abs() < 5 * torch.sqrt(pi * (1 - pi) / N)).all(), (freq, pi)
    for k in range(K):
        z = ((x[comp == k] - locs[k].double()) / math.exp(float(ls[k]))).numpy()
        z = z[np.abs(z) < 15]                                        # drop the rare cross-assigned tail points
        assert stats.kstest(z[:50000], "logistic").pvalue > 1e-4
    # --- reproducibility: torch.manual_seed controls the Philox key, successive calls advance the offset
    prm = B.DMoLParams(row.expand(1000, 3 * K).contiguous(), K, 1, -7.0)
    torch.manual_seed(7)
    a1 = lik.sample(B.DMoLParams(prm.raw, K, 1, -7.0))
    a2 = lik.sample(B.DMoLParams(prm.raw, K, 1, -7.0))
    torch.manual_seed(7)
    b1 = lik.sample(B.DMoLParams(prm.raw, K, 1, -7.0))
    assert torch.equal(a1, b1) and not torch.equal(a1, a2)
    # --- clamp: a wide component next to the upper edge
    wide = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.99, 0.99, 0.99, 0.99, -1.0, -1.0, -1.0, -1.0]).cuda()
    xs = lik.sample(B.DMoLParams(wide.expand(20000, 12).contiguous(), K, 1, -7.0))
    assert float(xs.max()) == 1.0 and float(xs.min()) >= -1.0 and (xs == 1.0).float().mean().item() > 0.3
    # --- bf16 parameters are read directly
    xs16 = lik.sample(B.DMoLParams(wide.expand(64, 12).contiguous().to(torch.bfloat16), K, 1, -7.0))
    assert xs16.dtype == torch.float32 and float(xs16.abs().max()) <= 1.0


@pytest.mark.parametrize("T,K", [(127, 10), (1000, 1), (333, 5), (130, 30), (64, 7), (2052, 1), (4096, 2), (1540, 5), (2048, 3)])
def test_no_out_of_bounds_writes(T, K, B):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are checked with canaries: every output of the
    DMoL / KL kernels lives inside one arena with poisoned guard bands on both sides (tail tiles, unaligned slabs,
    multi-sample tiles and the generic kernel are all covered by the shapes above)."""
    from blvm_b200 import ops
    lib = B._lib.lib
    dev, Bn, nb, G = "cuda", 3, 65536, 4096
    P = 3 * K
    y = torch.rand(Bn, T, device=dev) * 2 - 1
    raw = torch.randn(Bn, T, P, device=dev)
    x_sl = torch.tensor([T, T // 2, 1], device=dev)
    chunks = int(lib.blvm_dmol_chunks(T, K, 1))
    ev = lambda n: n + (n & 1)                                            # keep every view 8-byte aligned
    sizes = {"lp": ev(Bn * T), "graw": ev(Bn * T * P), "part": 2 * Bn * chunks}   # partials are fp64 = 2 floats each
    arena = torch.full((sum(sizes.values()) + G * (len(sizes) + 1),), 1234.5, device=dev)
    views, off = {}, G
    for name, n in sizes.items():
        views[name] = arena[off:off + n]
        off += n + G
    used = {"lp": Bn * T, "graw": Bn * T * P}
    part = views["part"].view(torch.float64)
    lp_v, graw_v = views["lp"][:used["lp"]], views["graw"][:used["graw"]]
    ops._dmol_call(y, raw, x_sl, None, -1e-3, Bn, T, K, 1, nb, -7.0, 1, lp_v.view(Bn, T), graw_v.view(Bn, T, P), part)
    torch.cuda.synchronize()
    off = 0
    for name, n in sizes.items():
        assert (arena[off:off + G] == 1234.5).all(), f"guard before {name} overwritten"
        off += G + n
    assert (arena[off:off + G] == 1234.5).all(), "trailing guard overwritten"
    assert torch.isfinite(lp_v).all() and torch.isfinite(graw_v).all() and torch.isfinite(part).all()
    assert (lp_v != 1234.5).all() and (graw_v != 1234.5).all()                     # every output element was written
    for name in used:                                                              # alignment padding untouched
        assert (views[name][used[name]:] == 1234.5).all()
    # KL kernel with a row length that is not a multiple of the vector width / tile
    Tz, Z = 37, 3
    ins = [torch.randn(Bn, Tz, Z, device=dev), torch.rand(Bn, Tz, Z, device=dev) + 0.1, torch.randn(Bn, Tz, Z, device=dev),
           torch.rand(Bn, Tz, Z, device=dev) + 0.1]
    n = Bn * Tz * Z
    kc = int(lib.blvm_kl_chunks(Tz * Z))
    npad = n + (n & 1)
    arena2 = torch.full((4 * npad + 4 * Bn * kc + 7 * G,), 77.0, device=dev)
    outs, off = [], G
    for _ in range(4):
        outs.append(arena2[off:off + n])
        off += npad + G
    pk = arena2[off:off + 2 * Bn * kc].view(torch.float64)
    off += 2 * Bn * kc + G
    pf = arena2[off:off + 2 * Bn * kc].view(torch.float64)
    lens = torch.tensor([Tz, 5, 0], device=dev)
    rc = lib.blvm_kl_elbo_fwd_grad(*[t.data_ptr() for t in ins], lens.data_ptr(), Bn, Tz, Z, 0.5, 1e-3, None,
                                   *[o.data_ptr() for o in outs], pk.data_ptr(), pf.data_ptr(), 0, ops._stream())
    assert rc == 0
    torch.cuda.synchronize()
    guard_mask = torch.ones_like(arena2, dtype=torch.bool)
    off = G
    for _ in range(4):
        guard_mask[off:off + n] = False
        off += npad + G
    guard_mask[off:off + 2 * Bn * kc] = False
    off += 2 * Bn * kc + G
    guard_mask[off:off + 2 * Bn * kc] = False
    assert (arena2[guard_mask] == 77.0).all(), "KL kernel wrote outside its outputs"
    assert all(torch.isfinite(o).all() for o in outs) and (outs[0].view(Bn, Tz, Z)[2] == 0).all()


# ---- sibling likelihoods behind the same boundary ---------------------------------------------------
@pytest.mark.parametrize("K", [1, 5, 10, 20, 7])
def test_gmm_module_golden(K, B, O):
    """DiagonalGaussianMixtureDense (`--likelihood GMM`): the sd activation + gaussian_mixture_ll + autograd of the
    reference (blvm/modules/distributions.py:153-204, blvm/utils/log_likelihoods.py:42-60) against the fused kernel."""
    g = load_golden(f"gmm_K{K}")
    lik = B.DiagonalGaussianMixtureDense(3 * K, 1, num_mix=K, initial_sd=1, epsilon=1e-4)
    assert lik.out_features == 3 * K and list(lik.state_dict().keys()) == ["params.weight", "params.bias"]
    assert math.isclose(lik.softplus_beta, float(g["beta"]))
    raw = cu(g["raw"]).requires_grad_(True)
    params = B.GMMParams(raw, K, 1, lik.softplus_beta, lik.epsilon)
    lp = lik.log_prob(cu(g["y"]), params)
    (lp * cu(g["gout"])).sum().backward()
    fac = gmm_row_factor(g["y"], g["raw"], K, float(g["beta"]), float(g["sd_add"]))
    assert_values_close(lp.detach().cpu().numpy(), g["lp64"], "GMM log-prob")
    assert_grads_close(raw.grad.cpu().numpy(), g["graw64"], K, np.abs(g["gout"]), "GMM grads", row_factor=fac)
    # the container indexes like the reference's (logits, mu, sd) tuple
    np.testing.assert_allclose(params[2].detach().cpu().numpy(), g["sd64"], rtol=1e-5)
    # functional form with separate, already activated tensors (sd given)
    t = [cu(g["raw"][:, :K]).requires_grad_(True), cu(g["raw"][:, K:2 * K]).unsqueeze(1).requires_grad_(True),
         cu(g["sd64"]).requires_grad_(True)]
    lp2 = B.gaussian_mixture_ll(cu(g["y"]), *t, epsilon=0)
    lp2.sum().backward()
    assert_values_close(lp2.detach().cpu().numpy(), g["lp64"], "GMM functional", rtol=2e-5)
    assert torch.isfinite(t[2].grad).all() and t[2].grad.abs().sum() > 0
    # fused ELBO with the GMM likelihood (WaveNet-style: no latents)
    Bn, T = 4, raw.shape[0] // 4
    x_sl = torch.tensor([T, T - 3, T // 2, 1])
    r = cu(g["raw"]).view(Bn, T, 3 * K).clone().requires_grad_(True)
    out = B.fused_elbo(cu(g["y"]).view(Bn, T), B.GMMParams(r, K, 1, lik.softplus_beta, lik.epsilon), x_sl, (), num_bins=2,
                       want_twise=True)
    out.loss.backward()
    m = O.sequence_mask(x_sl.numpy(), max_len=T)
    ref_rows = (g["lp64"].reshape(Bn, T) * m).sum(1)
    # this golden has sd down to 1e-4, i.e. single samples with |lp| ~ 1e7 that dominate their row: the row sum then carries
    # the fp32 rounding of ONE z^2 (~4e-7 relative, for the reference's fp32 run as well), so the 1e-6 bar of the DMoL sums
    # has no margin here and the last ulp of the mixture algebra decides; 4e-6 is 10 ulp of that term
    assert_sums_close(out.log_prob.cpu().numpy(), ref_rows, "GMM row sums", rtol=4e-6)
    assert_sums_close(out.loss.item(), -ref_rows.sum() / float(x_sl.sum()), "GMM loss", rtol=4e-6)
    assert (r.grad.cpu().numpy()[~m] == 0).all()


def test_gaussian_ll_and_mc_kl(B):
    g = load_golden("gaussian_ll")
    for e, eps in (("0", 0.0), ("1", 1e-2)):
        mu, sd = cu(g["mu_q"]).requires_grad_(True), cu(g["sd_q"]).requires_grad_(True)
        lp = B.gaussian_ll(cu(g["y"]), mu, sd, epsilon=eps, reduce_dim=None)
        (lp * cu(g["gout"])).sum().backward()
        assert_values_close(lp.detach().cpu().numpy(), g[f"lp64_{e}"], "gaussian_ll")
        ref = g[f"g_mu64_{e}"]
        np.testing.assert_allclose(mu.grad.cpu().numpy(), ref, rtol=RTOL, atol=1e-7 * np.abs(ref).max())
        if eps == 0:
            ref = g["g_sd64_0"]
            np.testing.assert_allclose(sd.grad.cpu().numpy(), ref, rtol=RTOL, atol=1e-7 * np.abs(ref).max())
        else:
            assert sd.grad is None or (sd.grad == 0).all()          # detached by the reference's no_grad clamp
    # reduce_dim = -1 sums the last axis like the reference (log_likelihoods.py:39)
    lp_sum = B.gaussian_ll(cu(g["y"]), cu(g["mu_q"]), cu(g["sd_q"]), epsilon=0)
    np.testing.assert_allclose(lp_sum.cpu().numpy(), g["lp64_0"].sum(-1), rtol=1e-5)
    ins = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
    kl = B.kl_divergence_gaussian_mc(*ins, cu(g["y"]))
    (kl * cu(g["gout"])).sum().backward()
    assert_values_close(kl.detach().cpu().numpy(), g["klmc64"], "MC KL", atol=2e-6)
    for t, n in zip(ins, ("mu_q", "sd_q", "mu_p", "sd_p")):
        ref = g[f"klmc_g_{n}64"]
        np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=RTOL, atol=1e-7 * np.abs(ref).max(), err_msg=n)


# ---- persistent pipelined kernel (dmol_stream_kernel) vs one tile per CTA (dmol_tile_kernel) ----------------------------
@pytest.mark.parametrize("K", [1, 2, 3, 5])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Bn,T", [(3, 4096), (7, 20000), (300, 1024), (2, 2052)])
def test_stream_kernel_bit_identical_to_tile_kernel(K, dtype, Bn, T, B):
    """Small-K shapes with 16-byte aligned slabs run the persistent TMA-ring kernel; it must reproduce the tile kernel
    bit for bit (values, gradients; the fp64 partial sums to 1e-13, see below) — ragged lengths, skipped padding tiles, upstream gradients,
    forward-only, more tiles than resident CTAs (300 x 1024) and a short tail tile (2052)."""
    from blvm_b200 import ops
    lib = B._lib.lib
    gen = torch.Generator().manual_seed(K * 1000 + T)
    nb = 65536
    y = (torch.randint(0, nb, (Bn, T), generator=gen).float() / (nb - 1) * 2 - 1).cuda()
    raw = torch.randn(Bn, T, 3 * K, generator=gen)
    raw[..., K:2 * K] = y.cpu().unsqueeze(-1) + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    raw = raw.to(dtype).cuda()
    x_sl = torch.randint(1, T + 1, (Bn,), generator=gen)
    x_sl[0] = T
    x_sl[-1] = 5                                       # most tiles of the last utterance lie in the padding
    x_dev = x_sl.cuda()
    gout = torch.randn(Bn, T, generator=gen).cuda()
    chunks = int(lib.blvm_dmol_chunks(T, K, 1))
    res = {}
    prev = lib.blvm_set_stream_mode(1)
    try:
        for mode in (0, 1):
            lib.blvm_set_stream_mode(mode)
            for flags in (1, 3):                       # mask output; + skip fully padded tiles
                for go in (None, gout):
                    lp = torch.full((Bn, T), 7.0, device="cuda")
                    graw = torch.full_like(raw, 7.0)
                    part = torch.full((Bn * chunks,), 7.0, dtype=torch.float64, device="cuda")
                    ops._dmol_call(y, raw, x_dev, go, -0.37, Bn, T, K, 1, nb, -7.0, flags, lp, graw, part)
                    lp2 = torch.full((Bn, T), 7.0, device="cuda")
                    part2 = torch.full((Bn * chunks,), 7.0, dtype=torch.float64, device="cuda")
                    ops._dmol_call(y, raw, x_dev, None, 0.0, Bn, T, K, 1, nb, -7.0, flags, lp2, None, part2)
                    torch.cuda.synchronize()
                    res[(mode, flags, go is not None)] = (lp, graw, part, lp2, part2)
    finally:
        lib.blvm_set_stream_mode(prev)
    for key, tile_out in res.items():
        if key[0] != 0:
            continue
        stream_out = res[(1,) + key[1:]]
        for a, b, name in zip(tile_out, stream_out, ("lp", "graw", "partials", "lp (fwd only)", "partials (fwd only)")):
            if name.startswith("partials"):
                # fp64 tile sums: the stream kernel runs 256-thread CTAs for fp32 K >= 3 (other warp grouping of the same 512 values)
                np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-13, atol=1e-9, err_msg=name)
            else:
                assert torch.equal(a, b), f"{name} differs between the tile and the stream kernel, flags={key[1]} gout={key[2]}"
        assert torch.equal(tile_out[0], tile_out[3])   # forward-only values equal the fwd+grad values
    # sanity against the oracle on one utterance (the stream path itself is what the golden tests exercise for K <= 5)
    lp = res[(1, 1, False)][0]
    assert torch.isfinite(lp).all() and (lp[-1, 5:] == 0).all() and (res[(1, 1, False)][1][-1, 5:] == 0).all()


def test_plain_backward_skips_fill_and_rescale(B):
    """`loss.backward()` on the fused op's loss hands autograd a persistent device-side 1.0: no rescale launch; a real
    upstream factor (GradScaler-style) still goes through the device-side rescale and gives factor x the same grads."""
    from blvm_b200 import ops
    g = load_golden("elbo_srnn_a")
    K, nb = int(g["K"]), int(g["num_bins"])

    def run(factor):
        raw = cu(g["raw"]).requires_grad_(True)
        kl = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
        out = B.fused_elbo(cu(g["y"]), B.DMoLParams(raw, K, 1, -7.0), torch.tensor(g["x_sl"]), [B.KLLevel(*kl, stride=int(g["stride"]))],
                           float(g["beta"]), float(g["free_nats"]), num_bins=nb)
        assert type(out.loss).__name__ == "FusedLoss" and type(out.loss * 2) is torch.Tensor
        ops.reset_launch_count()
        if factor is None:
            out.loss.backward()
        else:
            (out.loss * factor).backward()
        n = ops.launch_count()
        return n, raw.grad.clone(), [t.grad.clone() for t in kl], float(out.loss.detach())

    n0, graw0, gkl0, loss0 = run(None)
    n1, graw1, gkl1, _ = run(1.0)
    n2, graw2, gkl2, _ = run(1024.0)
    assert n0 == 0 and n1 == 1 and n2 == 1            # launches of OUR kernels inside backward
    assert torch.equal(graw0, graw1) and all(torch.equal(a, b) for a, b in zip(gkl0, gkl1))
    assert torch.equal(graw0 * 1024.0, graw2) and all(torch.equal(a * 1024.0, b) for a, b in zip(gkl0, gkl2))   # power of two: exact
    np.testing.assert_allclose(loss0, float(g["loss64"]), rtol=1e-6)
    scale = np.abs(g["graw64"]).max()
    np.testing.assert_allclose(graw0.cpu().numpy(), g["graw64"], rtol=1e-4, atol=1e-6 * scale)


def test_overlapped_launches_are_race_free(B, O):
    """The KL levels are launched as programmatic dependents of the likelihood kernel (they start while it drains and the
    finalize kernel waits for all of them): 40 repetitions of a 3-level step whose DMoL grid spans several waves must be
    bit-identical, and equal to the oracle."""
    gen = torch.Generator().manual_seed(77)
    Bn, T, K, nb = 24, 24576, 10, 65536
    strides, Zs = (64, 512, 4096), (32, 16, 8)
    y = (torch.randint(0, nb, (Bn, T), generator=gen).float() / (nb - 1) * 2 - 1)
    raw = torch.randn(Bn, T, 3 * K, generator=gen)
    raw[..., K:2 * K] = y.unsqueeze(-1) + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    x_sl = torch.randint(T // 2, T + 1, (Bn,), generator=gen)
    kls = []
    for s, Z in zip(strides, Zs):
        t = [torch.randn(Bn, T // s, Z, generator=gen) for _ in range(4)]
        t[1], t[3] = torch.nn.functional.softplus(t[1]) + 1e-3, torch.nn.functional.softplus(t[3]) + 1e-3
        kls.append(t)
    yd, rawd = y.cuda(), raw.cuda().requires_grad_(True)
    klsd = [[t.cuda().requires_grad_(True) for t in kl] for kl in kls]
    first = None
    for it in range(40):
        rawd.grad = None
        for kl in klsd:
            for t in kl:
                t.grad = None
        out = B.fused_elbo(yd, B.DMoLParams(rawd, K, 1, -7.0), x_sl, [B.KLLevel(*kl, stride=s, free_nats=0.25 * s / strides[0])
                                                                      for kl, s in zip(klsd, strides)], 0.7, 0.25, num_bins=nb)
        out.loss.backward()
        cur = [out.sums.clone(), out.log_prob.clone(), out.kl.clone(), out.kl_fn.clone(), rawd.grad.clone()] + [t.grad.clone() for kl in klsd for t in kl]
        if first is None:
            first = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(first, cur)), f"repetition {it} differs from the first"
    ref = O.fused_elbo_value_and_grad(y[:4].numpy(), raw[:4].numpy(), x_sl[:4].numpy(),
                                      [dict(mu_q=kl[0][:4].numpy(), sd_q=kl[1][:4].numpy(), mu_p=kl[2][:4].numpy(), sd_p=kl[3][:4].numpy(),
                                            stride=s, free_nats=0.25 * s / strides[0]) for kl, s in zip(kls, strides)], 0.7, K, nb)
    assert_sums_close(first[1][:4].cpu().numpy(), ref["logp"], "log p rows")
    assert_sums_close(first[2][:4].cpu().numpy(), ref["kl"], "KL rows")
    assert_sums_close(first[3][:4].cpu().numpy(), ref["kl_fn"], "free-nats KL rows")


@pytest.mark.parametrize("K", [1, 4, 5, 10, 30])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_sample_mode_tile_kernel_matches_direct_kernel(K, dtype, B):
    """The TMA-staged sample/mode kernel (aligned slabs) and the direct kernel (here forced by a 4-byte-offset view) draw
    the SAME samples for a given (seed, offset), and the categorical draw follows softmax(logits) when the logits vary
    per sample."""
    from blvm_b200 import ops
    lib = B._lib.lib
    gen = torch.Generator().manual_seed(900 + K)
    N = 128 * 64 + 48                                   # 16-byte multiple of bytes for every K; last tile is partial
    raw = torch.randn(N, 3 * K, generator=gen)
    raw[:, K:2 * K] = torch.linspace(-0.8, 0.8, K) if K > 1 else 0.0   # well separated components (>= 60 scales apart)
    raw[:, 2 * K:] = -8.0                                              # clamped to -7 inside: scale 9e-4
    raw = raw.to(dtype).cuda()
    esz = raw.element_size()
    buf = torch.empty(raw.numel() + 16 // esz, dtype=dtype, device="cuda")
    off = 4 // esz if esz == 4 else 2                   # 4-byte offset: not 16-byte aligned -> direct kernel
    shifted = buf[off:off + raw.numel()].view(N, 3 * K)
    shifted.copy_(raw)
    assert raw.data_ptr() % 16 == 0 and shifted.data_ptr() % 16 != 0
    outs = []
    for t in (raw, shifted):
        smp = torch.empty(N, 1, device="cuda")
        mode = torch.empty(N, 1, device="cuda")
        idx = torch.empty(N, dtype=torch.int32, device="cuda")
        rc = lib.blvm_dmol_sample_mode(t.data_ptr(), ops._DTYPE_CODE[dtype], N, K, 1, -7.0, 1234567, 42, smp.data_ptr(), mode.data_ptr(),
                                       idx.data_ptr(), ops._stream())
        assert rc == 0
        torch.cuda.synchronize()
        outs.append((smp, mode, idx))
    for a, b, name in zip(outs[0], outs[1], ("sample", "mode", "mode index")):
        assert torch.equal(a, b), f"{name}: tile kernel and direct kernel disagree"
    logits = raw[:, :K].float()
    assert torch.equal(torch.gather(logits, 1, outs[0][2].long().unsqueeze(1))[:, 0], logits.max(-1).values)   # an argmax (ties: bf16)
    assert torch.equal(outs[0][1][:, 0], torch.gather(raw[:, K:2 * K].float(), 1, outs[0][2].long().unsqueeze(1))[:, 0])
    assert float(outs[0][0].abs().max()) <= 1.0
    if K >= 4:
        # chosen component ~ softmax(logits): the mean log-probability of the drawn components matches its expectation
        comp = (outs[0][0] - raw[:, K:2 * K].float()).abs().argmin(-1)
        lsm = torch.log_softmax(logits.double(), -1)
        got = torch.gather(lsm, 1, comp.unsqueeze(1)).mean().item()
        want = (lsm.exp() * lsm).sum(-1).mean().item()
        assert abs(got - want) < 0.05, (got, want)      # N = 8240 draws: the standard error is ~0.015


def test_fp16_single_pass_with_known_grad_scaler(B):
    """`--use_amp True` (fp16 autocast + GradScaler): with the scaler known, the fp16 likelihood gradient is written in the
    forward pass pre-multiplied by the scaler's device-side scale -- and so are the KL gradients (the KL kernel reads the same
    device scalar) -- so `scaler.scale(loss).backward()` launches ONE rescale that exits at once; results are bit-identical to
    the deferred two-pass path (power-of-two scales), also for a second, different upstream factor."""
    from blvm_b200 import amp, ops
    g = load_golden("elbo_srnn_a")
    K, nb = int(g["K"]), int(g["num_bins"])
    scaler = torch.amp.GradScaler("cuda", init_scale=4096.0)
    scaler.scale(torch.zeros(1, device="cuda"))          # the scale tensor is created lazily by the first scale() call
    assert scaler._scale is not None

    def run(use_scaler, extra=1.0):
        raw = cu(g["raw"]).to(torch.float16).requires_grad_(True)
        kl = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
        ops.reset_launch_count()
        out = B.fused_elbo(cu(g["y"]), B.DMoLParams(raw, K, 1, -7.0), torch.tensor(g["x_sl"]), [B.KLLevel(*kl, stride=int(g["stride"]))],
                           float(g["beta"]), float(g["free_nats"]), num_bins=nb, denom=3.0e5, grad_scaler=scaler if use_scaler else None)
        n_fwd = ops.launch_count()
        (scaler.scale(out.loss) * extra).backward()
        return n_fwd, ops.launch_count() - n_fwd, raw.grad.clone(), [t.grad.clone() for t in kl], float(out.loss.detach())

    assert amp.active_grad_scaler(torch.device("cuda", torch.cuda.current_device())) is None   # nothing registered: explicit only
    fa, ba, graw_a, gkl_a, loss_a = run(True)
    fb, bb, graw_b, gkl_b, loss_b = run(False)
    assert (fa, fb) == (3, 3) and ba == 1 and bb == 2     # a: one early-exit rescale of all gradients; b: KL rescale + gradient kernel
    assert loss_a == loss_b
    assert graw_a.dtype == torch.float16 and torch.equal(graw_a, graw_b)
    assert all(torch.equal(x, y) for x, y in zip(gkl_a, gkl_b))
    assert (graw_a != 0).float().mean().item() > 0.2 and torch.isfinite(graw_a).all()   # the padding is zero, the rest survives fp16
    _, _, graw_c, gkl_c, _ = run(True, extra=0.5)         # upstream gradient != scale: the generic device-side ratio
    _, _, graw_d, gkl_d, _ = run(False, extra=0.5)
    np.testing.assert_allclose(graw_c.float().cpu().numpy(), graw_d.float().cpu().numpy(), rtol=2e-3, atol=1e-7)
    assert all(torch.equal(x, y) for x, y in zip(gkl_c, gkl_d))
    # registered scalers are picked up without the argument
    amp.register_grad_scaler(scaler)
    try:
        assert amp.active_grad_scaler(torch.device("cuda", torch.cuda.current_device())) is scaler
    finally:
        amp._scalers.discard(scaler)


def test_cuda_graph_replay_matches_eager(B):
    """bench.py times the step replayed from a CUDA graph (programmatic dependent launches are captured as graph edges):
    the replay must compute exactly what the eager call computes, also after the inputs have been rewritten in place."""
    gen = torch.Generator().manual_seed(31)
    Bn, T, K, nb = 16, 8192, 10, 65536
    strides, Zs = (64, 512), (32, 16)

    def inputs(seed):
        g = torch.Generator().manual_seed(seed)
        y = (torch.randint(0, nb, (Bn, T), generator=g).float() / (nb - 1) * 2 - 1)
        raw = torch.randn(Bn, T, 3 * K, generator=g)
        raw[..., K:2 * K] = y.unsqueeze(-1) + 0.1 * raw[..., K:2 * K]
        raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
        kls = []
        for s, Z in zip(strides, Zs):
            t = [torch.randn(Bn, T // s, Z, generator=g) for _ in range(4)]
            t[1], t[3] = torch.nn.functional.softplus(t[1]) + 1e-3, torch.nn.functional.softplus(t[3]) + 1e-3
            kls.append(t)
        return y, raw, kls

    x_sl = torch.randint(T // 2, T + 1, (Bn,), generator=gen)
    x_dev = x_sl.cuda()
    lens = [B.level_lengths(x_dev, s) for s in strides]
    denom = float(x_sl.sum())
    y0, raw0, kls0 = inputs(1)
    y_d, raw_d = y0.cuda(), raw0.cuda().requires_grad_(True)
    kl_d = [[t.cuda().requires_grad_(True) for t in kl] for kl in kls0]

    def step():
        raw_d.grad = None
        for kl in kl_d:
            for t in kl:
                t.grad = None
        out = B.fused_elbo(y_d, B.DMoLParams(raw_d, K, 1, -7.0), x_sl, [B.KLLevel(*kl, lens=ln) for kl, ln in zip(kl_d, lens)],
                           0.5, 0.0625, num_bins=nb, denom=denom, x_sl_device=x_dev)
        out.loss.backward()
        return out.sums, out.log_prob, raw_d.grad, [t.grad for kl in kl_d for t in kl]

    def snapshot(res):
        sums, logp, graw, gkl = res
        return [sums.clone(), logp.clone(), graw.clone()] + [g.clone() for g in gkl]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static = step()
    for seed in (1, 2, 3):
        y1, raw1, kls1 = inputs(seed)
        with torch.no_grad():
            y_d.copy_(y1)
            raw_d.copy_(raw1)
            for kl, new in zip(kl_d, kls1):
                for t, n in zip(kl, new):
                    t.copy_(n)
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        replayed = snapshot(static)
        eager = snapshot(step())
        torch.cuda.synchronize()
        for a, b, name in zip(replayed, eager, ["sums", "log_prob", "graw"] + [f"gkl{i}" for i in range(8)]):
            assert torch.equal(a, b), f"seed {seed}: {name} differs between graph replay and the eager call"

"""CPU check of the closed forms the CUDA kernels evaluate: csrc/blvm_math.cuh is compiled for the host (g++, libm in
place of the MUFU approximations) and compared with the reference's fp64 golden vectors under the same tolerances as
the GPU parity suite.  Catches algebra / branch / sign errors without a GPU; the last ulps are checked by -m gpu."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from parity import assert_grads_close, assert_values_close, gmm_row_factor

SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "hostsim", "_build")
LIB = os.path.join(OUT_DIR, "libhostsim.so")
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module", params=["libm", "mufu"])
def sim(request):
    """libm: exact transcendentals.  mufu: ex2 / lg2 / rcp results degraded to the error class of the MUFU approximations
    (2^-22 relative, 2^-22 absolute for lg2 near 0; blvm_math.cuh BLVM_HOSTSIM_MUFU_BITS) so that a closed form that only
    passes with libm's last ulps fails here rather than on the GPU."""
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = LIB if request.param == "libm" else LIB.replace(".so", "_mufu.so")
    extra = [] if request.param == "libm" else ["-DBLVM_HOSTSIM_MUFU_BITS=2"]
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", *extra, "-I",
                    os.path.join(ROOT, "benchmarking-lvms_b200", "csrc"), "-o", lib, SRC], check=True)
    return ctypes.CDLL(lib)


def P(a):
    return a.ctypes.data_as(FP)


def run_dmol(sim, g, force_generic=0):
    K, D, nb = int(g["K"]), int(g["D"]), int(g["num_bins"])
    y = np.ascontiguousarray(g["y"], np.float32)
    raw = np.ascontiguousarray(g["raw"], np.float32)
    gout = np.ascontiguousarray(g["gout"], np.float32)
    N = raw.shape[0]
    lp = np.empty(N, np.float32)
    gr = np.empty_like(raw)
    mode = sim.hostsim_dmol(P(y), P(raw), P(gout), ctypes.c_int64(N), K, D, nb, ctypes.c_float(-7.0), force_generic, P(lp), P(gr))
    if not force_generic and D == 1:
        assert mode == (1 if nb == 65536 else 0)  # 16-bit bins + the -7 clamp select the tiny-u specialisation
    return lp, gr


@pytest.mark.parametrize("case", ["dmol_K1_nb256", "dmol_K1_nb65536", "dmol_K2_nb256", "dmol_K2_nb65536", "dmol_K10_nb256",
                                  "dmol_K10_nb65536", "dmol_K30_nb256", "dmol_K30_nb65536", "dmol_K10_nb65536_D2"])
@pytest.mark.parametrize("force_generic", [0, 1])
def test_dmol_closed_forms(sim, case, force_generic):
    g = load_golden(case)
    lp, gr = run_dmol(sim, g, force_generic)
    assert_values_close(lp, g["lp64"], "log-prob")
    assert_grads_close(gr, g["graw64"], int(g["K"]), np.abs(g["gout"]), "grads")


def test_k1_two_samples_per_instruction_bit_identical(sim):
    """K = 1 in 16-bit mode pairs two samples per packed fp32x2 instruction inside the persistent kernel; the pairing
    (including pairs that contain an edge-bin sample, which fall back) must not change a single bit."""
    g = load_golden("dmol_K1_nb65536")
    lp_ref, gr_ref = run_dmol(sim, g)
    y = np.ascontiguousarray(g["y"], np.float32)
    raw = np.ascontiguousarray(g["raw"], np.float32)
    gout = np.ascontiguousarray(g["gout"], np.float32)
    N = raw.shape[0] - raw.shape[0] % 2
    lp = np.empty(N, np.float32)
    gr = np.empty((N, 3), np.float32)
    sim.hostsim_k1_pairs(P(y), P(raw), P(gout), ctypes.c_int64(N), 65536, ctypes.c_float(-7.0), P(lp), P(gr))
    assert np.array_equal(lp, lp_ref[:N]) and np.array_equal(gr, gr_ref[:N])


def test_dmol_non_power_of_two_bins(sim):
    """num_bins = 255: the y-edge thresholds are not exact in fp32, so the predicates follow the reference's fp32
    compare (rows where the fp64 run decides differently are compared with its fp32 run), and one forced row sits on
    the cdf_delta threshold where the reference's two arms differ by log(nb/(nb-1))."""
    from oracle import blvm_oracle as O
    g = load_golden("dmol_K5_nb255")
    K, nb = 5, 255
    lp, gr = run_dmol(sim, g)
    y32 = g["y"].astype(np.float32)
    e32 = np.where(y32 > np.float32(1 - 2 / nb), 2, np.where(y32 < np.float32(2 / nb - 1), 1, 0))
    y64 = y32.astype(np.float64)
    e64 = np.where(y64 > 1 - 2 / nb, 2, np.where(y64 < 2 / nb - 1, 1, 0))
    flip = (e32 != e64).any(-1)
    assert flip.sum() >= 4
    _, delta64 = O.dmol_branches(y64, g["raw"].astype(np.float64), K, 1, nb)
    knife = (np.abs(delta64 / 1e-5 - 1) < 1e-5).reshape(len(flip), -1).any(-1)
    ok = ~flip & ~knife
    assert_values_close(lp[ok], g["lp64"][ok], "log-prob")
    assert_grads_close(gr[ok], g["graw64"][ok], K, np.abs(g["gout"])[ok], "grads")
    np.testing.assert_allclose(lp[flip], g["lp32"][flip], rtol=2e-3)   # bit-exact predicate, fp32-noisy reference value
    assert (np.abs(lp[knife] - g["lp64"][knife]) <= np.log(nb / (nb - 1)) + 1e-4).all()


def test_dmol_small_m_gradients_do_not_cancel(sim):
    """y within 1e-4 scale units of a component mean: d/d loc ~ -inv * m / 2 must keep relative accuracy."""
    rng = np.random.default_rng(0)
    N, K = 4096, 1
    for nb in (256, 65536):
        y = (rng.integers(0, nb, N) / (nb - 1) * 2 - 1).astype(np.float32).clip(-0.99, 0.99)
        ls = rng.uniform(-6.5, -1.0, N).astype(np.float32)
        mu = (y + np.exp(ls) * rng.uniform(-3e-3, 3e-3, N)).astype(np.float32)
        raw = np.stack([np.zeros(N, np.float32), mu, ls], -1)
        g = dict(K=1, D=1, num_bins=nb, y=y[:, None], raw=raw, gout=np.ones(N, np.float32))
        _, gr = run_dmol(sim, g)
        from oracle import blvm_oracle as O
        _, ref = O.dmol_value_and_grad(y.astype(np.float64), raw.astype(np.float64), 1, 1, nb)
        m = (y.astype(np.float64) - mu) * np.exp(-ls.astype(np.float64))
        sel = np.abs(m) > 1e-5  # below that the fp32 rounding of (y - mu) itself dominates
        rel = np.abs(gr[sel, 1] - ref[sel, 1]) / np.abs(ref[sel, 1])
        assert rel.max() < 5e-5, rel.max()


def test_dl_closed_forms(sim):
    for nb in (256, 65536):
        g = load_golden(f"dl_nb{nb}")
        N = g["raw"].shape[0]
        lp = np.empty(N, np.float32)
        gr = np.empty((N, 2), np.float32)
        sim.hostsim_dl(P(np.ascontiguousarray(g["y"])), P(np.ascontiguousarray(g["raw"])), P(np.ascontiguousarray(g["gout"])),
                       ctypes.c_int64(N), nb, ctypes.c_float(-7.0), P(lp), P(gr))
        assert_values_close(lp, g["lp64"], "DL log-prob")
        assert_grads_close(gr, g["graw64"], 1, np.abs(g["gout"]), "DL grads")


def test_kl_closed_forms(sim):
    g = load_golden("kl_free_nats")
    ins = [np.ascontiguousarray(g[n]) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
    n = ins[0].size
    Z = ins[0].shape[-1]
    gout = np.ascontiguousarray(g["gout"])
    for i, fn in enumerate(g["free_nats"]):
        outs = [np.empty(ins[0].shape, np.float32) for _ in range(6)]
        sim.hostsim_kl(*[P(a) for a in ins], P(gout), ctypes.c_int64(n), ctypes.c_float(fn / Z), int(fn != 0),
                       *[P(o) for o in outs])
        assert_values_close(outs[0], g["kl64"], "kl", atol=1e-6)
        tie = np.zeros(ins[0].shape, bool)
        if i == 3:
            tie[tuple(g["tie_index"])] = True  # an exact fp32 tie of the reference's formula; not a tie in fp64
        assert_values_close(outs[1][~tie], g[f"klfn64_{i}"][~tie], "kl_fn", atol=1e-6)
        for o, nme in zip(outs[2:], ("mu_q", "sd_q", "mu_p", "sd_p")):
            ref = g[f"g_{nme}64_{i}"]
            np.testing.assert_allclose(o[~tie], ref[~tie], rtol=1e-5, atol=1e-7 * np.abs(ref).max(), err_msg=nme)
    assert (outs[0][0, 0, :4] == 0).all()  # q == p gives exactly 0


@pytest.mark.parametrize("K", [1, 5, 10, 20, 7])
@pytest.mark.parametrize("force_generic", [0, 1])
def test_gmm_closed_forms(sim, K, force_generic):
    """Gaussian mixture from the packed Linear output (softplus + epsilon inside, chain rule folded in) and from
    (logits, mu, sd) against the reference's fp64 run."""
    from oracle import blvm_oracle as O
    g = load_golden(f"gmm_K{K}")
    y = np.ascontiguousarray(g["y"][:, 0], np.float32)
    raw = np.ascontiguousarray(g["raw"], np.float32)
    gout = np.ascontiguousarray(g["gout"], np.float32)
    N = raw.shape[0]
    lp = np.empty(N, np.float32)
    gr = np.empty_like(raw)
    sim.hostsim_gmm(P(y), P(raw), P(gout), ctypes.c_int64(N), K, ctypes.c_float(float(g["beta"])), ctypes.c_float(float(g["sd_add"])),
                    ctypes.c_float(0.0), 1, force_generic, P(lp), P(gr))
    assert_values_close(lp, g["lp64"], "GMM log-prob (from raw)")
    assert_grads_close(gr, g["graw64"], K, np.abs(g["gout"]), "GMM grads (from raw)",
                       row_factor=gmm_row_factor(g["y"], g["raw"], K, float(g["beta"]), float(g["sd_add"])))
    # the oracle's closed form agrees with the reference's autograd too
    L, G = O.gmm_value_and_grad(g["y"].astype(np.float64), g["raw"].astype(np.float64), K, 1, float(g["beta"]), float(g["sd_add"]),
                                g["gout"].astype(np.float64))
    np.testing.assert_allclose(L, g["lp64"], rtol=1e-12)
    assert np.abs(G - g["graw64"]).max() <= 1e-9 * np.abs(g["graw64"]).max()
    # "sd given" variant: pack [logits | mu | sd] with the reference's own fp64 sd
    packed = np.ascontiguousarray(np.concatenate([g["raw"][:, :2 * K], g["sd64"][:, 0].astype(np.float32)], -1), np.float32)
    sim.hostsim_gmm(P(y), P(packed), P(gout), ctypes.c_int64(N), K, ctypes.c_float(1.0), ctypes.c_float(0.0), ctypes.c_float(0.0),
                    0, force_generic, P(lp), P(gr))
    assert_values_close(lp, g["lp64"], "GMM log-prob (sd given)", rtol=2e-5)   # sd itself was rounded to fp32


def test_gaussian_ll_closed_forms(sim):
    from oracle import blvm_oracle as O
    g = load_golden("gaussian_ll")
    y, mu, sd, gout = (np.ascontiguousarray(g[k], np.float32).reshape(-1) for k in ("y", "mu_q", "sd_q", "gout"))
    n = y.size
    for e, eps in (("0", 0.0), ("1", 1e-2)):
        lp, gm, gs = (np.empty(n, np.float32) for _ in range(3))
        sim.hostsim_gauss(P(y), P(mu), P(sd), P(gout), ctypes.c_int64(n), ctypes.c_float(eps), P(lp), P(gm), P(gs))
        assert_values_close(lp, g[f"lp64_{e}"].reshape(-1), "gaussian_ll")
        np.testing.assert_allclose(gm, g[f"g_mu64_{e}"].reshape(-1), rtol=1e-5, atol=1e-7 * np.abs(g[f"g_mu64_{e}"]).max())
        np.testing.assert_allclose(gs, g[f"g_sd64_{e}"].reshape(-1), rtol=1e-5, atol=1e-7 * np.abs(g[f"g_sd64_0"]).max())
        ol, om, os_ = O.gaussian_ll_value_and_grad(g["y"].astype(np.float64), g["mu_q"].astype(np.float64), g["sd_q"].astype(np.float64),
                                                   eps, g["gout"].astype(np.float64))
        np.testing.assert_allclose(ol, g[f"lp64_{e}"], rtol=1e-12)
        np.testing.assert_allclose(om, g[f"g_mu64_{e}"], rtol=1e-10)
        np.testing.assert_allclose(os_, g[f"g_sd64_{e}"], rtol=1e-10, atol=1e-300)
    assert (g["g_sd64_1"] == 0).all()          # the reference's no_grad clamp detaches sd (kept)
    kl = O.kl_divergence_gaussian_mc(*[g[k].astype(np.float64) for k in ("mu_q", "sd_q", "mu_p", "sd_p", "y")])
    np.testing.assert_allclose(kl, g["klmc64"], rtol=1e-12)


@pytest.mark.parametrize("nb", [65536, 256])
@pytest.mark.parametrize("K", [2, 5, 10, 30])
@pytest.mark.parametrize("force_generic", [0, 1])
def test_dmol_extreme_regimes_against_fp64_oracle(sim, nb, K, force_generic):
    """Beyond the goldens: logits up to +-100, locations 1e-4 .. 3 away from the target in either direction, log-scales
    from -12 (clamped) to +4, edge-bin targets — the regimes where a closed form could overflow, cancel or lose the
    absolute precision of a log-prob close to 0.  The fp32 device math stays within the parity tolerances of the
    reference run in fp64 (the reference's own fp32 run is off by up to 80x that tolerance there)."""
    from oracle import blvm_oracle as O
    rng = np.random.default_rng(1000 + K + nb)
    N = 2000
    y = (rng.integers(0, nb, N) / (nb - 1) * 2 - 1).astype(np.float32)
    y[:40] = rng.choice(np.array([-1.0, 1.0, 2 / nb - 1, 1 - 2 / nb], np.float32), 40)
    raw = np.empty((N, 3 * K), np.float32)
    raw[:, :K] = rng.normal(0, 1, (N, K)) * rng.choice([1, 10, 40], (N, 1))
    raw[:, K:2 * K] = y[:, None] + rng.normal(0, 1, (N, K)) * rng.choice([1e-4, 1e-2, 0.3, 3.0], (N, K))
    raw[:, 2 * K:] = rng.uniform(-12, 4, (N, K))
    gout = rng.normal(0, 1, N).astype(np.float32)
    lp = np.empty(N, np.float32)
    gr = np.empty_like(raw)
    sim.hostsim_dmol(P(y), P(raw), P(gout), ctypes.c_int64(N), K, 1, nb, ctypes.c_float(-7.0), force_generic, P(lp), P(gr))
    assert np.isfinite(lp).all() and np.isfinite(gr).all()
    L, G = O.dmol_value_and_grad(y.astype(np.float64), raw.astype(np.float64), K, 1, nb, -7.0, gout.astype(np.float64))
    _, delta = O.dmol_branches(y.astype(np.float64).reshape(N, 1), raw.astype(np.float64), K, 1, nb, -7.0)   # (N, 1, K)
    ok = ~(np.abs(delta / 1e-5 - 1) < 2e-4).any(axis=(1, 2))     # knife-edge rows: the reference's two arms differ there
    assert_values_close(lp[ok], L[ok], "log-prob (extreme regimes)")
    assert_grads_close(gr[ok], G[ok], K, np.abs(gout[ok]), "grads (extreme regimes)")


def test_kl_extreme_regimes_against_fp64_oracle(sim):
    """Standard deviations from e^-12 to e^5, ratios sd_q/sd_p from 1 +- 1e-6 to e^+-5, mean gaps from 0 to 30: the KL stays
    within the parity tolerance of the fp64 reference and every gradient within 1e-6 of the magnitude of its (possibly
    cancelling) terms."""
    from oracle import blvm_oracle as O
    rng = np.random.default_rng(3)
    n = 100000
    mu_q = (rng.normal(0, 1, n) * rng.choice([1e-3, 1, 30], n)).astype(np.float32)
    mu_p = (mu_q + rng.normal(0, 1, n) * rng.choice([0, 1e-6, 1e-3, 1, 30], n)).astype(np.float32)
    sd_q = np.exp(rng.uniform(-12, 5, n)).astype(np.float32)
    sd_p = (sd_q * np.exp(rng.normal(0, 1, n) * rng.choice([0, 1e-6, 1e-3, 1, 5], n))).astype(np.float32)
    outs = [np.empty(n, np.float32) for _ in range(6)]
    sim.hostsim_kl(P(mu_q), P(sd_q), P(mu_p), P(sd_p), None, ctypes.c_int64(n), ctypes.c_float(0.0), 0, *[P(o) for o in outs])
    assert all(np.isfinite(o).all() for o in outs)
    q, sq, p, sp = (a.astype(np.float64) for a in (mu_q, sd_q, mu_p, sd_p))
    kl, _, grads = O.kl_value_and_grad(q, sq, p, sp, free_nats=0.0, gout=None)
    assert_values_close(outs[0], kl, "KL (extreme regimes)")
    d2 = (q - p) ** 2
    scales = (np.abs(q - p) / sp ** 2, 1 / sq + sq / sp ** 2, np.abs(q - p) / sp ** 2, 1 / sp + (sq ** 2 + d2) / sp ** 3)
    for o, g, sc, nm in zip(outs[2:], grads, scales, ("mu_q", "sd_q", "mu_p", "sd_p")):
        assert (np.abs(o - g) <= 1e-6 * sc + 1e-30).all(), nm


def test_kl_mc_closed_forms(sim):
    """Monte-Carlo KL log q(z) - log p(z) (variational.py:73-83) and its gradients against the reference's fp64 run."""
    g = load_golden("gaussian_ll")
    z, mu_q, sd_q, mu_p, sd_p, gout = (np.ascontiguousarray(g[k], np.float32).reshape(-1) for k in ("y", "mu_q", "sd_q", "mu_p", "sd_p", "gout"))
    n = z.size
    outs = [np.empty(n, np.float32) for _ in range(6)]
    sim.hostsim_kl_mc(P(z), P(mu_q), P(sd_q), P(mu_p), P(sd_p), P(gout), ctypes.c_int64(n), *[P(o) for o in outs])
    ref = g["klmc64"].reshape(-1)
    # a difference of two log-densities: the error is bounded relative to the size of its terms (the golden contains
    # sd_q = 1e-3 rows with |kl| ~ 1e6), 1e-6 of them
    aq, ap = (z - mu_q) / sd_q, (z - mu_p) / sd_p
    scale = 0.5 * aq.astype(np.float64) ** 2 + 0.5 * ap.astype(np.float64) ** 2 + np.abs(np.log(sd_q.astype(np.float64) / sd_p))
    assert (np.abs(outs[0] - ref) <= 1e-6 * scale + 1e-7).all()
    for o, nm in zip(outs[1:5], ("mu_q", "sd_q", "mu_p", "sd_p")):
        r = g[f"klmc_g_{nm}64"].reshape(-1)
        np.testing.assert_allclose(o, r, rtol=1e-5, atol=1e-7 * np.abs(r).max(), err_msg=nm)
    # d/dz is not in the golden: finite differences of the fp64 oracle
    from oracle import blvm_oracle as O
    z64, q, sq, p, sp = (a.astype(np.float64) for a in (z, mu_q, sd_q, mu_p, sd_p))
    h = 1e-6
    fd = (O.kl_divergence_gaussian_mc(q, sq, p, sp, z64 + h) - O.kl_divergence_gaussian_mc(q, sq, p, sp, z64 - h)) / (2 * h) * gout
    np.testing.assert_allclose(outs[5], fd, rtol=2e-4, atol=1e-5 * np.abs(fd).max())


def test_linear_domain_agrees_with_log_domain_and_falls_back():
    """The 16-bit kernels evaluate the mixture in the linear domain (blvm_math.cuh: dmol_sample_lin) with a per-sample fallback to
    the log-domain body.  Host builds with and without it: on ordinary samples the two agree to a few ulp but are NOT the same
    computation (the linear path is really taken); on samples far from every component, on edge bins and on non-finite parameters
    the results are bit-identical (the fallback IS the log-domain body)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    sims = {}
    for tag, flag in (("lin", "-DBLVM_LINEAR_DOMAIN=1"), ("log", "-DBLVM_LINEAR_DOMAIN=0")):
        lib = LIB.replace(".so", f"_{tag}.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", flag, "-I",
                        os.path.join(ROOT, "benchmarking-lvms_b200", "csrc"), "-o", lib, SRC], check=True)
        sims[tag] = ctypes.CDLL(lib)
    g = load_golden("dmol_K10_nb65536")
    lp_lin, gr_lin = run_dmol(sims["lin"], g)
    lp_log, gr_log = run_dmol(sims["log"], g)
    assert_values_close(lp_lin, lp_log.astype(np.float64), "linear vs log domain", rtol=2e-6)
    assert (lp_lin != lp_log).mean() > 0.05, "the linear-domain path does not seem to be compiled in"
    assert_values_close(lp_lin, g["lp64"], "linear domain vs the reference in fp64")
    assert_grads_close(gr_lin, g["graw64"], 10, np.abs(g["gout"]), "linear-domain grads")
    # far-out / edge / non-finite samples: the fallback
    rng = np.random.default_rng(3)
    K, nb, N = 10, 65536, 4000
    y = (rng.integers(0, nb, N) / (nb - 1) * 2 - 1).astype(np.float32)
    y[:4] = np.array([-1.0, 1.0, -1.0, 1.0], np.float32)
    raw = rng.normal(size=(N, 3 * K)).astype(np.float32)
    raw[:, K:2 * K] = y[:, None] + rng.choice([-1.0, 1.0], size=(N, K)) * rng.uniform(0.2, 1.5, (N, K))   # >= 0.2 away ...
    raw[:, 2 * K:] = rng.uniform(-9.0, -5.7, (N, K))                                                       # ... at scales <= 3.3e-3
    raw[10, 0] = np.nan
    raw[11, 2 * K] = np.inf
    gout = rng.normal(size=N).astype(np.float32)
    fake = dict(K=K, D=1, num_bins=nb, y=y.reshape(N, 1), raw=raw, gout=gout)
    a_lp, a_gr = run_dmol(sims["lin"], fake)
    b_lp, b_gr = run_dmol(sims["log"], fake)
    assert np.array_equal(a_lp, b_lp, equal_nan=True) and np.array_equal(a_gr, b_gr, equal_nan=True)
    assert np.isfinite(a_lp[20:]).all() and float(np.abs(a_lp[20:]).min()) > 50.0       # the log-domain values are finite and huge

"""Pin the C oracle (oracle/blvm_oracle.c — also the timed CPU baseline of bench.py) against the golden vectors
generated from the reference, in fp64 (tight) and fp32 (the reference's native precision; loose where the reference's
own fp32 arithmetic is ill-conditioned)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import c_oracle as C


@pytest.mark.parametrize("name", ["elbo_vrnn_a", "elbo_srnn_a", "elbo_vrnn_b"])
def test_c_oracle_elbo_fp64(name):
    g = load_golden(name)
    lv = [dict(mu_q=g["mu_q"].astype(np.float64), sd_q=g["sd_q"].astype(np.float64), mu_p=g["mu_p"].astype(np.float64),
               sd_p=g["sd_p"].astype(np.float64), stride=int(g["stride"]), free_nats=float(g["free_nats"]))]
    r = C.elbo_step(g["y"].astype(np.float64), g["raw"].astype(np.float64), g["x_sl"], lv, float(g["beta"]), int(g["K"]),
                    int(g["num_bins"]))
    np.testing.assert_allclose(r["loss"], g["loss64"], rtol=1e-12)
    np.testing.assert_allclose(r["elbo"], g["elbo64"], rtol=1e-12)
    np.testing.assert_allclose(r["logp"], g["logp64"], rtol=1e-12)
    assert np.abs(r["graw"] - g["graw64"]).max() / np.abs(g["graw64"]).max() < 1e-8
    for i, nme in enumerate(("mu_q", "sd_q", "mu_p", "sd_p")):
        ref = g[f"g_{nme}64"]
        assert np.abs(r["gkl"][0][i] - ref).max() <= 1e-10 * np.abs(ref).max()


def test_c_oracle_elbo_fp32_tracks_reference_fp32():
    g = load_golden("elbo_srnn_a")
    lv = [dict(mu_q=g["mu_q"], sd_q=g["sd_q"], mu_p=g["mu_p"], sd_p=g["sd_p"], stride=int(g["stride"]),
               free_nats=float(g["free_nats"]))]
    r = C.elbo_step(g["y"], g["raw"], g["x_sl"], lv, float(g["beta"]), int(g["K"]), int(g["num_bins"]))
    np.testing.assert_allclose(r["loss"], g["loss32"], rtol=2e-5)
    np.testing.assert_allclose(r["elbo"], g["elbo32"], rtol=2e-5)


@pytest.mark.parametrize("case", ["dmol_K1_nb256", "dmol_K10_nb65536", "dmol_K30_nb256", "dmol_K2_nb65536"])
def test_c_oracle_dmol_per_sample(case):
    g = load_golden(case)
    K, nb = int(g["K"]), int(g["num_bins"])
    N = g["raw"].shape[0]
    y = g["y"].reshape(1, N).astype(np.float64)
    raw = g["raw"].reshape(1, N, 3 * K).astype(np.float64)
    lp, graw, rows = C.dmol(y, raw, np.array([N]), K, nb)
    np.testing.assert_allclose(lp[0], g["lp64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(rows[0], g["lp64"].sum(), rtol=1e-12)
    # gradient for gout == 1 equals the golden gradient divided by gout
    gout = g["gout"].astype(np.float64)[:, None]
    sc = np.abs(g["graw64"] / gout).max(-1, keepdims=True) + 1e-30
    assert (np.abs(graw[0] - g["graw64"] / gout) / sc).max() < 1e-6
    assert C.num_threads() >= 1

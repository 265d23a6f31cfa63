"""Tolerances of the parity suite, written once.

north_star: per-sample log-probs, KLs and gradients within fp32 rel 1e-5 of the reference; ELBO sums within rel 1e-6;
predicates / bin indices bit-exact.  "The reference" for floating point is the reference's own code run in fp64 on the
same fp32 inputs (tests/golden/*64 arrays): its fp32 run deviates from that by up to 1e-3 relative on 16-bit audio
because `sigmoid(a) - sigmoid(b)` cancels, so fp32-vs-fp32 at 1e-5 is not a meaningful bar — where
the fp32 reference agrees with its fp64 self to 1e-5 we are additionally checked against it directly.

Gradient entries are compared relative to the scale of their parameter group within the sample (logits / locs /
log-scales): components with negligible responsibility have gradients ~1e-20 whose individual relative error is
meaningless in fp32 (for the reference as well).
"""
import numpy as np

RTOL = 1e-5        # per-sample values and gradients
ATOL_LP = 1e-6     # log-probs are O(1..20) nats; absolute floor for values that approach 0
RTOL_SUM = 1e-6    # per-utterance sums / loss


def assert_values_close(ours, ref64, what="value", rtol=RTOL, atol=ATOL_LP, mask=None):
    ours = np.asarray(ours, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    err = np.abs(ours - ref64)
    tol = rtol * np.abs(ref64) + atol
    bad = err > tol
    if mask is not None:
        bad &= mask
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} outside rtol={rtol} atol={atol}; worst err/tol = "
                           f"{(err / tol)[bad].max():.3g} at {np.argwhere(bad)[0]}")


def gmm_row_factor(y, raw, K, beta, sd_add):
    """fp32 conditioning of a Gaussian-mixture gradient: responsibilities are exp(lp_k - max) and lp_k carries an
    absolute rounding error of ~|lp_k| * 6e-8 in fp32 (for the reference as well), so their relative accuracy is bounded
    by that; with sd down to 1e-4, |lp_k| reaches 1e7.  Returns max(1, max_k |lp_k| * 1.2e-7 / RTOL) per sample."""
    y = np.asarray(y, np.float64).reshape(-1, 1)
    raw = np.asarray(raw, np.float64)
    mu, p = raw[:, K:2 * K], raw[:, 2 * K:]
    with np.errstate(over="ignore"):
        sd = np.where(p * beta > 20, p, np.log1p(np.exp(np.minimum(beta * p, 20))) / beta) + sd_add
    lp = -0.5 * ((y - mu) / sd) ** 2 - np.log(sd) - 0.9189385332046727
    return np.maximum(1.0, np.abs(lp).max(-1) * 1.2e-7 / RTOL)


def assert_grads_close(ours, ref64, K, gout_abs, what="grad", rtol=RTOL, rows=None, row_factor=None):
    """ours/ref64 (N, P) with P = K(2D+1) laid out [logits K | per d: locs K, log-scales K]; gout_abs (N,) = |upstream
    gradient| of each sample (sets the absolute floor: d/d logit is bounded by it)."""
    ours = np.asarray(ours, dtype=np.float64).reshape(-1, np.asarray(ref64).shape[-1])
    ref64 = np.asarray(ref64, dtype=np.float64).reshape(ours.shape)
    gout_abs = np.asarray(gout_abs, dtype=np.float64).reshape(-1, 1)
    P = ours.shape[-1]
    worst = 0.0
    for g0 in range(0, P, K):
        o, r = ours[:, g0:g0 + K], ref64[:, g0:g0 + K]
        gmax = np.abs(r).max(-1, keepdims=True)
        rt = rtol if row_factor is None else rtol * np.asarray(row_factor, dtype=np.float64).reshape(-1, 1)
        tol = rt * np.abs(r) + rt * gmax + 1e-6 * gout_abs + 1e-30
        ratio = np.abs(o - r) / tol
        if rows is not None:
            ratio = ratio[rows]
        worst = max(worst, float(ratio.max()) if ratio.size else 0.0)
    assert worst <= 1.0, f"{what}: worst err/tol = {worst:.3g} (rtol={rtol} of the per-sample group scale)"
    return worst


def assert_sums_close(ours, ref64, what="sum", rtol=RTOL_SUM):
    ours = np.asarray(ours, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    np.testing.assert_allclose(ours, ref64, rtol=rtol, atol=1e-9, err_msg=what)

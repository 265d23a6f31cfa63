"""CPU oracle (numpy) for the blvm DMoL + Gaussian-KL + masked-ELBO path.  TEST INFRASTRUCTURE ONLY.

This file restates, op by op, the reference algorithm of JakobHavtorn/benchmarking-lvms for the one hot path this
repo accelerates.  It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  Nothing under `benchmarking-lvms_b200/` (the product) imports
it, and the product has no CPU fallback.

Parity pin: the reference's own tests hold no fixture for this path, so the oracle is pinned against
outputs of the reference itself, generated in the build container by `tests/golden/make_golden.py` (fp32 and fp64
runs of the unmodified `/root/reference/blvm` functions) and committed under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against them.

All functions are dtype-generic: they compute in the dtype of their inputs (np.float32 reproduces the reference's
native arithmetic up to libm ulps, np.float64 is "the truth" used for tolerances).  Python-float constants follow
numpy's weak-scalar rule, i.e. they are rounded to the array dtype exactly like torch rounds Python scalars.

Reference citations are `path:line` under /root/reference.
"""
import math

import numpy as np

__all__ = [
    "sequence_mask", "discretized_logistic_ll", "discretized_logistic_mixture_ll", "split_dmol_params",
    "dmol_branches", "dmol_value_and_grad", "dl_value_and_grad", "kl_divergence_gaussian", "kl_value_and_grad",
    "discount_free_nats", "elbo_vrnn", "elbo_srnn", "elbo_cwvae", "elbo_stcn", "loss_wavenet", "quantize", "bits_per_dim", "fused_elbo_value_and_grad",
    "softplus_beta", "gaussian_ll", "gaussian_mixture_ll", "gmm_value_and_grad", "gaussian_ll_value_and_grad",
    "kl_divergence_gaussian_mc",
]


# ---------------------------------------------------------------------------------------------------------------
# small torch-op restatements
# ---------------------------------------------------------------------------------------------------------------
def _sigmoid(x):
    with np.errstate(over="ignore"):
        return 1 / (1 + np.exp(-x))


def _softplus(x):
    """F.softplus with beta=1, threshold=20: x if x > 20 else log1p(exp(x))."""
    with np.errstate(over="ignore"):
        return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def _logsumexp(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0)
    return (np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m).squeeze(axis)


def _log_softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    s = x - m
    return s - np.log(np.sum(np.exp(s), axis=axis, keepdims=True))


def sequence_mask(seq_lens, stride=1, max_len=None, dtype=bool):
    """blvm/utils/operations.py:90-119: arange(T)[None, :] < seq_lens[:, None]; T = max_len or ceil(max/stride)."""
    seq_lens = np.asarray(seq_lens)
    T = max_len or math.ceil(seq_lens.max() / stride)
    return (np.arange(T)[None, :] < seq_lens[:, None]).astype(dtype)


# ---------------------------------------------------------------------------------------------------------------
# discretized logistic (mixture) log-likelihood
# ---------------------------------------------------------------------------------------------------------------
def _dl_terms(y, loc, log_scale, num_bins):
    """The shared per-element body of log_likelihoods.py:133-165 and :198-227. Returns (log_prob, branch, aux)."""
    centered_y = y - loc                                              # :134 / :202
    inv_stdv = np.exp(-log_scale)                                     # :135 / :203
    plus_in = inv_stdv * (centered_y + 1.0 / (num_bins - 1))          # :138 / :206  half width 1/(nb-1)
    cdf_plus = _sigmoid(plus_in)
    minus_in = inv_stdv * (centered_y - 1.0 / (num_bins - 1))         # :140 / :208
    cdf_minus = _sigmoid(minus_in)
    cdf_delta = cdf_plus - cdf_minus                                  # :142 / :210
    log_cdf_plus = plus_in - _softplus(plus_in)                       # :145 / :213
    log_one_minus_cdf_minus = -_softplus(minus_in)                    # :148 / :216
    mid_in = inv_stdv * centered_y                                    # :151 / :219
    log_pdf_mid = mid_in - log_scale - 2.0 * _softplus(mid_in)        # :152 / :220
    with np.errstate(divide="ignore"):
        big = cdf_delta > 1e-5
        safe = np.where(big, np.log(np.maximum(cdf_delta, 1e-10)), log_pdf_mid - math.log(num_bins / 2))  # :153 / :221
    lower = y < 2 / num_bins - 1                                      # :158 / :226   edges at 2/nb (other convention)
    upper = y > 1 - 2 / num_bins                                      # :159 / :227
    log_prob = np.where(lower, log_cdf_plus, safe)
    log_prob = np.where(upper, log_one_minus_cdf_minus, log_prob)
    # branch id: 0 lower edge, 1 upper edge, 2 cdf_delta, 3 mid-pdf fallback (upper wins over lower like the where-chain)
    branch = np.where(upper, 1, np.where(lower, 0, np.where(big, 2, 3))).astype(np.int8)
    aux = dict(inv=inv_stdv, a=plus_in, b=minus_in, m=mid_in, sa=cdf_plus, sb=cdf_minus, delta=cdf_delta)
    return log_prob, branch, aux


def discretized_logistic_ll(y, loc, log_scale, num_bins=256, reduce_dim=-1):
    """blvm/utils/log_likelihoods.py:98-166. `reduce_dim` falsy => no reduction (:166)."""
    y = np.asarray(y)
    assert y.max() <= 1.0 and y.min() >= -1.0                         # :131
    y, loc, log_scale = np.broadcast_arrays(y, loc, log_scale)
    log_prob, _, _ = _dl_terms(y, loc, log_scale, num_bins)
    if reduce_dim:
        return log_prob.squeeze(reduce_dim) if log_prob.shape[reduce_dim] == 1 else log_prob.sum(reduce_dim)  # :10-14
    return log_prob


def discretized_logistic_mixture_ll(y, logit_probs, locs, log_scales, num_bins=256, reduce_dim=-1):
    """blvm/utils/log_likelihoods.py:170-231. y (*, D); logit_probs (*, K); locs, log_scales (*, D, K) -> (*)."""
    y = np.asarray(y)
    assert y.max() <= 1.0 and y.min() >= -1.0                         # :195
    y = y[..., None]                                                  # :198-199 (expand over the mixture dim)
    y, locs, log_scales = np.broadcast_arrays(y, locs, log_scales)
    log_prob, _, _ = _dl_terms(y, locs, log_scales, num_bins)         # (*, D, K)
    ax = reduce_dim - 1
    log_prob = log_prob.squeeze(ax) if log_prob.shape[ax] == 1 else log_prob.sum(ax)  # :229
    log_prob = log_prob + _log_softmax(np.asarray(logit_probs), -1)   # :230
    return _logsumexp(log_prob, -1)                                   # :231


def split_dmol_params(raw, K, D=1, log_epsilon=-7.0):
    """blvm/modules/distributions.py:383-387: raw (*, K(2D+1)) -> logits (*,K), locs (*,D,K), clamped log_scales."""
    raw = np.asarray(raw)
    logit_probs = raw[..., :K]
    lls = raw[..., K:].reshape(*raw.shape[:-1], D, 2 * K)
    locs, log_scales = lls[..., :K], lls[..., K:]
    return logit_probs, locs, np.maximum(log_scales, log_epsilon), log_scales


def dmol_branches(y, raw, K, D=1, num_bins=256, log_epsilon=-7.0):
    """Branch id per (sample, d, k) and cdf_delta — diagnostics for the parity tests."""
    _, locs, ls, _ = split_dmol_params(raw, K, D, log_epsilon)
    yb, locs, ls = np.broadcast_arrays(np.asarray(y)[..., None], locs, ls)
    _, branch, aux = _dl_terms(yb, locs, ls, num_bins)
    return branch, aux["delta"]


def _dl_grad_terms(branch, aux):
    """d log_prob / d(plus_in, minus_in, mid_in) and the direct d/d log_scale, per branch.

    What torch autograd produces for log_likelihoods.py:213-227 (`where` routes the gradient to the selected branch
    only): lower edge d/da = 1 - sigmoid(a); upper edge d/db = -sigmoid(b); cdf_delta branch d/da = s_a(1-s_a)/delta,
    d/db = -s_b(1-s_b)/delta; fallback d/dm = 1 - 2 sigmoid(m), direct d/dls = -1 (SURVEY.md §8a closed forms).
    """
    sa, sb, delta, m = aux["sa"], aux["sb"], aux["delta"], aux["m"]
    zero = np.zeros_like(sa)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        da = np.where(branch == 0, 1 - sa, np.where(branch == 2, sa * (1 - sa) / delta, zero))
        db = np.where(branch == 1, -sb, np.where(branch == 2, -sb * (1 - sb) / delta, zero))
    dm = np.where(branch == 3, 1 - 2 * _sigmoid(m), zero)
    dls_direct = np.where(branch == 3, -np.ones_like(sa), zero)
    return da, db, dm, dls_direct


def dmol_value_and_grad(y, raw, K, D=1, num_bins=256, log_epsilon=-7.0, gout=None):
    """Per-sample DMoL log-prob from the packed Linear output and d(sum gout*lp)/d raw (closed form).

    y (N, D) or (N,), raw (N, K(2D+1)).  Follows distributions.py:383-387 + log_likelihoods.py:198-231; gradient =
    what `loss.backward()` gives through those lines, including the clamp's pass-at-equality rule (raw_ls >= eps).
    """
    raw = np.asarray(raw)
    y = np.asarray(y).reshape(raw.shape[0], D)
    logits, locs, ls, raw_ls = split_dmol_params(raw, K, D, log_epsilon)
    yb = np.broadcast_to(y[..., None], locs.shape)
    lp_dk, branch, aux = _dl_terms(yb, locs, ls, num_bins)            # (N, D, K)
    lp_k = lp_dk.sum(-2)                                              # (N, K)
    w = _log_softmax(logits, -1)
    v = lp_k + w
    L = _logsumexp(v, -1)                                             # (N,)
    if gout is None:
        gout = np.ones_like(L)
    gout = np.asarray(gout).astype(L.dtype)
    r = np.exp(v - L[..., None])                                      # posterior responsibilities (N, K)
    pi = np.exp(w)
    da, db, dm, dls_direct = _dl_grad_terms(branch, aux)
    inv, a, b, m = aux["inv"], aux["a"], aux["b"], aux["m"]
    dmu = -inv * (da + db + dm)                                       # d lp_dk / d loc
    dls = -(a * da + b * db + m * dm) + dls_direct                    # d lp_dk / d log_scale (clamped)
    dls = dls * (raw_ls >= log_epsilon)                               # clamp(min) passes gradient at equality
    g = gout[..., None]
    grad = np.empty_like(raw)
    grad[..., :K] = g * (r - pi)
    glls = grad[..., K:].reshape(raw.shape[0], D, 2 * K)
    glls[..., :K] = (g * r)[:, None, :] * dmu
    glls[..., K:] = (g * r)[:, None, :] * dls
    grad[..., K:] = glls.reshape(raw.shape[0], -1)
    return L, grad


def dl_value_and_grad(y, raw, num_bins=256, log_epsilon=-7.0, gout=None):
    """Single discretized logistic from the packed (N, 2) = [mu | log_scale] Linear output (distributions.py:303-307)."""
    raw = np.asarray(raw)
    y = np.asarray(y).reshape(-1)
    mu, raw_ls = raw[:, 0], raw[:, 1]
    ls = np.maximum(raw_ls, log_epsilon)
    lp, branch, aux = _dl_terms(y, mu, ls, num_bins)
    if gout is None:
        gout = np.ones_like(lp)
    da, db, dm, dls_direct = _dl_grad_terms(branch, aux)
    grad = np.empty_like(raw)
    grad[:, 0] = gout * (-aux["inv"] * (da + db + dm))
    grad[:, 1] = gout * (-(aux["a"] * da + aux["b"] * db + aux["m"] * dm) + dls_direct) * (raw_ls >= log_epsilon)
    return lp, grad


# ---------------------------------------------------------------------------------------------------------------
# sibling likelihoods: Gaussian and Gaussian mixture
# ---------------------------------------------------------------------------------------------------------------
def softplus_beta(x, beta, threshold=20.0):
    """nn.Softplus(beta): log1p(exp(beta x))/beta, linear above beta*x > threshold."""
    with np.errstate(over="ignore"):
        return np.where(x * beta > threshold, x, np.log1p(np.exp(np.minimum(beta * x, threshold))) / beta)


def gaussian_ll(y, mu, sd, epsilon=1e-6, reduce_dim=-1):
    """blvm/utils/log_likelihoods.py:17-39 (sd clamped at epsilon; the clamp is under no_grad in the reference)."""
    if epsilon:
        sd = np.maximum(sd, epsilon)
    log_prob = -((y - mu) ** 2) / (2 * sd ** 2) - np.log(sd) - 0.5 * math.log(2 * math.pi)
    if reduce_dim:
        return log_prob.squeeze(reduce_dim) if log_prob.shape[reduce_dim] == 1 else log_prob.sum(reduce_dim)
    return log_prob


def gaussian_mixture_ll(y, logits, mu, sd, epsilon=1e-6, reduce_dim=-1):
    """blvm/utils/log_likelihoods.py:42-60. y (*, D); logits (*, K); mu, sd (*, D, K) -> (*)."""
    log_prob_y = gaussian_ll(np.asarray(y)[..., None], mu, sd, epsilon=epsilon, reduce_dim=reduce_dim - 1)
    return _logsumexp(log_prob_y + _log_softmax(np.asarray(logits), -1), -1)


def gmm_value_and_grad(y, raw, K, D=1, beta=math.log(2), sd_add=1e-4, gout=None):
    """GMM log-prob from the packed Linear output (DiagonalGaussianMixtureDense.forward + log_prob,
    blvm/modules/distributions.py:189-204) and d(sum gout*lp)/d raw.  sd = softplus_beta(log_sd) + sd_add."""
    raw = np.asarray(raw)
    y = np.asarray(y).reshape(raw.shape[0], D)
    logits = raw[..., :K]
    mls = raw[..., K:].reshape(raw.shape[0], D, 2 * K)
    mu, p = mls[..., :K], mls[..., K:]
    sd = softplus_beta(p, beta) + sd_add
    z = (y[..., None] - mu) / sd
    lp_dk = -0.5 * z * z - np.log(sd) - 0.5 * math.log(2 * math.pi)
    v = lp_dk.sum(-2) + _log_softmax(logits, -1)
    L = _logsumexp(v, -1)
    if gout is None:
        gout = np.ones_like(L)
    r = np.exp(v - L[..., None])
    pi = np.exp(_log_softmax(logits, -1))
    with np.errstate(over="ignore"):
        dsd_dp = np.where(p * beta > 20.0, 1.0, 1 / (1 + np.exp(-beta * p)))
    g = np.asarray(gout)[..., None]
    grad = np.empty_like(raw)
    grad[..., :K] = g * (r - pi)
    gm = grad[..., K:].reshape(raw.shape[0], D, 2 * K)
    gm[..., :K] = (g * r)[:, None, :] * (z / sd)
    gm[..., K:] = (g * r)[:, None, :] * ((z * z - 1) / sd) * dsd_dp
    grad[..., K:] = gm.reshape(raw.shape[0], -1)
    return L, grad


def gaussian_ll_value_and_grad(y, mu, sd, epsilon=0.0, gout=None):
    """Elementwise gaussian_ll with gradients w.r.t. mu and sd (no gradient reaches sd when epsilon != 0: the reference
    clamps under no_grad, which detaches it, log_likelihoods.py:33-35)."""
    sd_c = np.maximum(sd, epsilon) if epsilon else sd
    z = (y - mu) / sd_c
    lp = -0.5 * z * z - np.log(sd_c) - 0.5 * math.log(2 * math.pi)
    if gout is None:
        gout = np.ones_like(lp)
    g_mu = gout * z / sd_c
    g_sd = gout * (z * z - 1) / sd_c * (0.0 if epsilon else 1.0)
    return lp, g_mu, g_sd


def kl_divergence_gaussian_mc(mu_q, sd_q, mu_p, sd_p, z, epsilon=0):
    """blvm/utils/variational.py:73-83: log q(z) - log p(z), elementwise."""
    return gaussian_ll(z, mu_q, sd_q, epsilon, None) - gaussian_ll(z, mu_p, sd_p, epsilon, None)


# ---------------------------------------------------------------------------------------------------------------
# Gaussian KL and free nats
# ---------------------------------------------------------------------------------------------------------------
def kl_divergence_gaussian(mu_q, sd_q, mu_p, sd_p):
    """blvm/utils/variational.py:67-70 (inputs are standard deviations)."""
    return np.log(sd_p) - np.log(sd_q) + (sd_q ** 2 + (mu_q - mu_p) ** 2) / (2 * sd_p ** 2) - 0.5


def discount_free_nats(kld, free_nats=None, shared_dims=None):
    """blvm/utils/variational.py:86-122: max(kld, free_nats / prod(shape[shared_dims])); identity if None/0."""
    if free_nats is None or free_nats == 0:
        return kld
    if isinstance(shared_dims, int):
        shared_dims = (shared_dims,)
    if shared_dims is not None:
        min_kl = free_nats / math.prod([kld.shape[d] for d in shared_dims])
    else:
        min_kl = free_nats
    return np.maximum(kld, kld.dtype.type(min_kl))


def kl_value_and_grad(mu_q, sd_q, mu_p, sd_p, free_nats=0.0, gout=None):
    """kl, kl_fn (shared_dims=-1) and d(sum gout*kl_fn)/d inputs.  torch.maximum splits the gradient 1/2-1/2 at
    exact ties."""
    kl = kl_divergence_gaussian(mu_q, sd_q, mu_p, sd_p)
    kl_fn = discount_free_nats(kl, free_nats, -1)
    if gout is None:
        gout = np.ones_like(kl)
    if free_nats is None or free_nats == 0:
        sel = np.ones_like(kl)
    else:
        c = kl.dtype.type(free_nats / kl.shape[-1])
        sel = np.where(kl > c, 1.0, np.where(kl == c, 0.5, 0.0)).astype(kl.dtype)
    g = gout * sel
    d = mu_q - mu_p
    g_mu_q = g * d / sd_p ** 2
    g_sd_q = g * (-1 / sd_q + sd_q / sd_p ** 2)
    g_sd_p = g * (1 / sd_p - (sd_q ** 2 + d ** 2) / sd_p ** 3)
    return kl, kl_fn, (g_mu_q, g_sd_q, -g_mu_q, g_sd_p)


# ---------------------------------------------------------------------------------------------------------------
# per-model ELBO reducers
# ---------------------------------------------------------------------------------------------------------------
def _dmol_lp_bt(y, raw, K, num_bins, log_epsilon=-7.0):
    logits, locs, ls, _ = split_dmol_params(raw, K, 1, log_epsilon)
    return discretized_logistic_mixture_ll(np.asarray(y)[..., None], logits, locs, ls, num_bins)  # (B, T)


def _elbo_vrnn_like(y, raw, kld_twise, x_sl, stride, beta, free_nats, K, num_bins, return_fn_kl):
    x_sl = np.asarray(x_sl)
    seq_mask = sequence_mask(x_sl, dtype=np.float64)                  # vrnn.py:266 dtype=float => float64
    T = seq_mask.shape[1]
    log_prob_twise = _dmol_lp_bt(y[:, :T], raw[:, :T], K, num_bins) * seq_mask   # :268
    log_prob = log_prob_twise.reshape(log_prob_twise.shape[0], -1).sum(1)        # :269
    seq_mask_kl = seq_mask[:, ::stride][..., None]                    # :271
    kld = (kld_twise * seq_mask_kl).sum((1, 2))                       # :272
    elbo = log_prob - kld                                             # :273
    kld_fn = (discount_free_nats(kld_twise, free_nats, -1) * seq_mask_kl).sum((1, 2))  # :275-276
    loss = -(log_prob - beta * kld_fn).sum() / x_sl.sum()             # :277
    return loss, elbo, log_prob, (kld_fn if return_fn_kl else kld), seq_mask


def elbo_vrnn(y, raw, kld_twise, x_sl, stride, beta=1, free_nats=0, K=10, num_bins=65536):
    """blvm/models/vrnn.py:255-279. Quirk kept: the returned `kld` is the free-nats-discounted one (:276-279)."""
    return _elbo_vrnn_like(y, raw, kld_twise, x_sl, stride, beta, free_nats, K, num_bins, True)


def elbo_srnn(y, raw, kld_twise, x_sl, stride, beta=1, free_nats=0, K=10, num_bins=65536):
    """blvm/models/srnn.py:137-160. Returns the raw KL (:153,160)."""
    return _elbo_vrnn_like(y, raw, kld_twise, x_sl, stride, beta, free_nats, K, num_bins, False)


def elbo_cwvae(y, raw, kld_layerwise, x_sl, overall_strides, beta=1, free_nats=0, K=10, num_bins=65536):
    """blvm/models/clockwork_vae/clockwork_vae.py:132-161 with the masks of :231-240 (bool)."""
    x_sl = np.asarray(x_sl)
    T = y.shape[1]
    seq_mask = sequence_mask(x_sl, max_len=T)
    log_prob_twise = _dmol_lp_bt(y, raw, K, num_bins) * seq_mask      # :143
    log_prob = log_prob_twise.reshape(y.shape[0], -1).sum(1)          # :144
    kld_l, klds_fn = [], []
    for l, kl in enumerate(kld_layerwise):
        lens = np.ceil(x_sl / overall_strides[l]).astype(np.int64)    # :237
        mask = sequence_mask(lens, max_len=kl.shape[1])[..., None]    # :238
        fn = free_nats * overall_strides[l] / overall_strides[0]      # :151
        kld_l.append((kl * mask).sum((1, 2)))                         # :152
        klds_fn.append((discount_free_nats(kl, fn, -1) * mask).sum((1, 2)))  # :153
    kld, kld_fn = sum(kld_l), sum(klds_fn)                            # :155
    elbo = log_prob - kld                                             # :157
    loss = -(log_prob - beta * kld_fn).sum() / x_sl.sum()             # :159
    return loss, elbo, log_prob, kld, kld_l


def elbo_stcn(y, raw, kl_inputs, x_sl, n_stack_frames, beta, free_nats, K=10, num_bins=65536, z=None):
    """blvm/models/stcn/stcn.py:256-297. kl_inputs = [(mu_q, sd_q, mu_p, sd_p)] per latent level; `z` (list, one per
    level) selects the bottom-up variant: Monte-Carlo KL log q(z) - log p(z) (:288) instead of the analytic KL (:286)."""
    x_sl = np.asarray(x_sl)
    log_prob_twise = _dmol_lp_bt(y, raw, K, num_bins)                 # :278
    seq_mask = sequence_mask(x_sl)                                    # :280
    log_prob = (log_prob_twise * seq_mask).sum(1)                     # :281
    z_mask = seq_mask[:, ::n_stack_frames][..., None]                 # :284
    if z is None:
        klds = [kl_divergence_gaussian(*ins) * z_mask for ins in kl_inputs]           # :286
    else:
        klds = [kl_divergence_gaussian_mc(*ins, z[l]) * z_mask for l, ins in enumerate(kl_inputs)]   # :288
    klds_fn = [discount_free_nats(k, free_nats, -1) * z_mask for k in klds]           # :289 (mask, fn, mask)
    kld = np.concatenate(klds, -1).sum((1, 2))                        # :290
    kld_fn = np.concatenate(klds_fn, -1).sum((1, 2))                  # :291
    klds = [k.sum((1, 2)) for k in klds]                              # :292
    elbo = log_prob - kld                                             # :295
    loss = -(log_prob - beta * kld_fn).sum() / x_sl.sum()             # :297
    return loss, elbo, log_prob, kld, klds


def loss_wavenet(y, raw, x_sl, K=10, num_bins=65536):
    """blvm/models/wavenet/wavenet.py:128-146."""
    x_sl = np.asarray(x_sl)
    seq_mask = sequence_mask(x_sl, max_len=y.shape[1])                # :141
    log_prob_twise = _dmol_lp_bt(y, raw, K, num_bins) * seq_mask      # :142
    log_prob = log_prob_twise.reshape(y.shape[0], -1).sum(1)          # :143
    loss = -np.nansum(log_prob) / np.nansum(x_sl)                     # :145
    return loss, log_prob, log_prob_twise


def fused_elbo_value_and_grad(y, raw, x_sl, kl_levels, beta, K, num_bins=65536, log_epsilon=-7.0, dtype=np.float64):
    """The whole path in one call, the shape the fused CUDA op has: values + d loss / d inputs.

    kl_levels: list of dict(mu_q, sd_q, mu_p, sd_p, stride, free_nats[, lens]) — per level the valid latent steps are
    `lens` if given else ceil(x_sl / stride) (equal to the `mask[:, ::stride]` of vrnn.py:271 / stcn.py:284 and to the
    level masks of clockwork_vae.py:237-238).  loss = -sum_b(logp_b - beta * kl_fn_b) / sum(x_sl).
    Returns dict(loss, logp, kl, kl_fn, elbo, lp_twise, graw, gkl=[(4 grads)]).
    """
    x_sl = np.asarray(x_sl)
    y = np.asarray(y, dtype=dtype)
    raw = np.asarray(raw, dtype=dtype)
    B, T = y.shape
    mask = sequence_mask(x_sl, max_len=T)
    denom = dtype(x_sl.sum())
    gout = (-(mask.astype(dtype)) / denom).reshape(-1)
    lp, graw = dmol_value_and_grad(y.reshape(-1), raw.reshape(B * T, -1), K, 1, num_bins, log_epsilon, gout)
    lp = lp.reshape(B, T) * mask
    logp = lp.sum(1)
    kl_tot = np.zeros(B, dtype)
    klfn_tot = np.zeros(B, dtype)
    gkl, kl_l = [], []
    for lv in kl_levels:
        ins = [np.asarray(lv[n], dtype=dtype) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
        lens = np.asarray(lv["lens"]) if lv.get("lens") is not None else np.ceil(x_sl / lv["stride"]).astype(np.int64)
        m = sequence_mask(lens, max_len=ins[0].shape[1])[..., None].astype(dtype)
        kl, kl_fn, grads = kl_value_and_grad(*ins, free_nats=lv["free_nats"], gout=m * (beta / denom))
        kl_l.append((kl * m).sum((1, 2)))
        kl_tot += kl_l[-1]
        klfn_tot += (kl_fn * m).sum((1, 2))
        gkl.append(grads)
    loss = -(logp - beta * klfn_tot).sum() / denom
    return dict(loss=loss, logp=logp, kl=kl_tot, kl_fn=klfn_tot, elbo=logp - kl_tot, lp_twise=lp,
                graw=graw.reshape(raw.shape), gkl=gkl, kl_levels=kl_l)


# ---------------------------------------------------------------------------------------------------------------
# integer side: Quantize, and the bits-per-dim arithmetic
# ---------------------------------------------------------------------------------------------------------------
def quantize(x, boundaries):
    """torch.bucketize(x, boundaries, right=False) (transforms.py:257): first index i with boundaries[i] >= x.
    `boundaries` is the fp32 table `torch.linspace(-1, 1, bins)` of transforms.py:249; ATen's vectorised linspace is
    not restated here (its lane-wise rounding is an implementation detail) — the golden file stores the table and the
    product builds it with torch.linspace on the host exactly like the reference does."""
    return np.searchsorted(boundaries, np.asarray(x, dtype=np.float32), side="left").astype(np.int64)


def bits_per_dim(elbo, x_sl):
    """blvm/evaluation/metrics.py:443-468 + :241-247: sum_b(-elbo_b / ln 2) / sum_b x_sl_b."""
    return float((-np.asarray(elbo) / math.log(2)).sum() / np.asarray(x_sl).sum())

"""Eager-PyTorch restatement of the reference path (the op chain the reference actually launches).  TEST INFRASTRUCTURE ONLY.

`blvm_oracle.py` restates the algorithm in numpy with closed-form gradients; this file restates it as the SAME SEQUENCE
OF TORCH OPS the reference executes (one eager kernel per op, backward through torch autograd), because that — not a C
loop — is what runs when a blvm experiment trains on a GPU.  Two uses, both on the checker side:
  * `tests/test_oracle_golden.py` pins it (values + autograd gradients, CPU) against the reference's goldens;
  * `bench.py --impl reference --reference-device cuda` times it on the GPU box, where /root/reference does not exist:
    the "eager PyTorch on the same B200" number of SURVEY.md §8(d), reported next to the CPU arm.
Nothing under `benchmarking-lvms_b200/` imports it.  Citations are `path:line` under /root/reference.
"""
import math

import torch
import torch.nn.functional as F


def sequence_mask(x_sl, max_len, dtype, device):
    """blvm/utils/operations.py:90-119."""
    return (torch.arange(max_len, device=device)[None, :] < x_sl.to(device)[:, None]).to(dtype)


def dmol_ll(y, logit_probs, locs, log_scales, num_bins):
    """blvm/utils/log_likelihoods.py:198-231 op for op.  y (*, D), logit_probs (*, K), locs / log_scales (*, D, K)."""
    y = y.unsqueeze(-1)                                                       # :198-199
    centered = y - locs                                                       # :202
    inv_stdv = torch.exp(-log_scales)                                         # :203
    half = 1.0 / (num_bins - 1)
    plus_in = inv_stdv * (centered + half)                                    # :206
    cdf_plus = torch.sigmoid(plus_in)                                         # :207
    minus_in = inv_stdv * (centered - half)                                   # :208
    cdf_minus = torch.sigmoid(minus_in)                                       # :209
    cdf_delta = cdf_plus - cdf_minus                                          # :210
    log_cdf_plus = plus_in - F.softplus(plus_in)                              # :213
    log_one_minus_cdf_minus = -F.softplus(minus_in)                           # :216
    mid_in = inv_stdv * centered                                              # :219
    log_pdf_mid = mid_in - log_scales - 2.0 * F.softplus(mid_in)              # :220
    inner = torch.where(cdf_delta > 1e-5, torch.log(torch.clamp(cdf_delta, min=1e-10)),
                        log_pdf_mid - math.log(num_bins / 2))                 # :221-223
    lp = torch.where(y < 2 / num_bins - 1, log_cdf_plus, inner)               # :226
    lp = torch.where(y > 1 - 2 / num_bins, log_one_minus_cdf_minus, lp)       # :227
    lp = lp.sum(-2) + torch.log_softmax(logit_probs, dim=-1)                  # :229-230
    return torch.logsumexp(lp, dim=-1)                                        # :231


def split_params(raw, K, D=1, log_epsilon=-7.0):
    """blvm/modules/distributions.py:383-387."""
    logit_probs = raw[..., :K]
    locs, log_scales = raw[..., K:].reshape(*raw.shape[:-1], D, 2 * K).chunk(2, dim=-1)
    return logit_probs, locs, log_scales.clamp(min=log_epsilon)


def kl_gaussian(mu_q, sd_q, mu_p, sd_p):
    """blvm/utils/variational.py:67-70."""
    return torch.log(sd_p) - torch.log(sd_q) + (sd_q ** 2 + (mu_q - mu_p) ** 2) / (2 * sd_p ** 2) - 0.5


def free_nats_max(kld, free_nats):
    """blvm/utils/variational.py:86-122 with shared_dims=-1."""
    if not free_nats:
        return kld
    return torch.maximum(kld, torch.tensor(free_nats / kld.shape[-1], dtype=kld.dtype, device=kld.device))


def elbo_step(y, raw, x_sl, kl_levels, beta, K, num_bins, mask_dtype=torch.float32, check_range=True):
    """One forward of the path with the structure of the reference's compute_elbo (vrnn.py:255-279 /
    clockwork_vae.py:132-161): masked DMoL sums, masked KL sums per level, free nats, loss.  `kl_levels` = list of
    (mu_q, sd_q, mu_p, sd_p, stride, free_nats).  `mask_dtype` float64 reproduces VRNN/SRNN (vrnn.py:266), float32
    stands in for the bool masks of CW-VAE/STCN/WaveNet.  Returns (loss, elbo, log_prob, kld)."""
    B, T = y.shape
    if check_range:
        assert torch.max(y) <= 1.0 and torch.min(y) >= -1.0                  # log_likelihoods.py:195 (a device->host sync)
    mask = sequence_mask(x_sl, T, mask_dtype, y.device)
    lp_twise = dmol_ll(y.unsqueeze(-1), *split_params(raw, K), num_bins) * mask
    log_prob = lp_twise.flatten(1).sum(1)
    kld = torch.zeros_like(log_prob)
    kld_fn = torch.zeros_like(log_prob)
    for mu_q, sd_q, mu_p, sd_p, stride, free_nats in kl_levels:
        kl = kl_gaussian(mu_q, sd_q, mu_p, sd_p)
        lens = torch.ceil(x_sl / stride).long()
        m = sequence_mask(lens, kl.shape[1], mask_dtype, y.device).unsqueeze(-1)
        kld = kld + (kl * m).sum((1, 2))
        kld_fn = kld_fn + (free_nats_max(kl, free_nats) * m).sum((1, 2))
    elbo = log_prob - kld
    loss = -(log_prob - beta * kld_fn).sum() / x_sl.sum().to(y.device)
    return loss, elbo, log_prob, kld

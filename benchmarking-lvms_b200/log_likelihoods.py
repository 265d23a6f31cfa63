"""Functional log-likelihoods with the reference's names and signatures (blvm/utils/log_likelihoods.py), backed by the
sm_100a kernels.  Only the discretized-logistic family is on the hot path (a2, a3)."""
import math
from typing import Optional

import torch

from . import ops

__all__ = ["discretized_logistic_mixture_ll", "discretized_logistic_ll", "gaussian_ll", "gaussian_mixture_ll", "reduce"]

_NO_CLAMP = -math.inf  # log-scales handed to the functional API are used as they are (the module clamps, not the function)


def reduce(tensor: torch.Tensor, dim, operation=torch.sum):
    """Same contract as blvm/utils/log_likelihoods.py:10-14."""
    if tensor.size(dim) == 1:
        return tensor.squeeze(dim)
    return operation(tensor, dim)


def _packed_source(logit_probs, locs, log_scales):
    """If the three tensors are the views our DiscretizedLogisticMixtureDense.forward produced, return its packed
    (raw, K, D, log_epsilon) so the kernel can read the Linear output directly (and clamp inside)."""
    src = getattr(log_scales, "_blvm_packed", None)
    if src is None:
        return None
    raw, K, D, log_eps = src
    if getattr(logit_probs, "_blvm_packed", (None,))[0] is raw and getattr(locs, "_blvm_packed", (None,))[0] is raw:
        return raw, K, D, log_eps
    return None


def discretized_logistic_mixture_ll(
    y: torch.Tensor,
    logit_probs: torch.Tensor,
    locs: torch.Tensor,
    log_scales: torch.Tensor,
    num_bins: int = 256,
    reduce_dim: int = -1,
):
    """Log-likelihood of a mixture of discretized logistics — drop-in for blvm/utils/log_likelihoods.py:170-231.

    Args (as in the reference): y (*, D); logit_probs (*, K); locs, log_scales (*, D, K), broadcastable against y
    (experiments/experiment_distribution_audio.py:122-131 passes (K,) / (1, K) parameters). Returns (*).

    One fused kernel evaluates the bin-edge sigmoids, the edge/mid/fallback cases, log_softmax and logsumexp; the
    autograd backward is a second launch of the same kernel in value+gradient mode.  `reduce_dim` must address the D
    axis of y (-1, the only value the reference's call sites use).  The range assert of :195 is deferred to
    `blvm_b200.check_input_range()` (no per-call device sync).
    """
    if reduce_dim not in (-1, y.ndim - 1):
        raise NotImplementedError("blvm_b200.discretized_logistic_mixture_ll reduces over the last (D) axis of y only")
    packed = _packed_source(logit_probs, locs, log_scales)
    if packed is not None and y.shape[:-1] == packed[0].shape[:-1]:
        raw, K, D, log_eps = packed
        return ops.dmol_log_prob(y, raw, K, D, num_bins, log_eps)

    # generic route: broadcast, pack [logits | per d: locs, log_scales] and run the same kernel without a clamp
    K = logit_probs.size(-1)
    if locs.ndim < 2 or log_scales.ndim < 2:
        locs = locs.reshape(*([1] * (2 - locs.ndim)), *locs.shape)
        log_scales = log_scales.reshape(*([1] * (2 - log_scales.ndim)), *log_scales.shape)
    D = y.size(-1)
    batch = torch.broadcast_shapes(y.shape[:-1], logit_probs.shape[:-1], locs.shape[:-2], log_scales.shape[:-2])
    logit_b = logit_probs.expand(*batch, K)
    locs_b = locs.expand(*batch, D, K)
    ls_b = log_scales.expand(*batch, D, K)
    raw = torch.cat([logit_b, torch.cat([locs_b, ls_b], dim=-1).flatten(-2)], dim=-1)  # (*, K(2D+1))
    y_b = y.expand(*batch, D)
    return ops.dmol_log_prob(y_b, raw, K, D, num_bins, _NO_CLAMP)


def discretized_logistic_ll(y: torch.Tensor, loc: torch.Tensor, log_scale: torch.Tensor, num_bins: int = 256,
                            reduce_dim: Optional[int] = -1):
    """Elementwise discretized-logistic log-likelihood — drop-in for blvm/utils/log_likelihoods.py:98-166.
    All dimensions independent; `reduce_dim` falsy => no reduction (:166)."""
    packed = getattr(log_scale, "_blvm_packed_dl", None)
    if packed is not None and getattr(loc, "_blvm_packed_dl", (None,))[0] is packed[0] and y.shape == loc.shape \
            and loc.shape[-1] == 1:
        raw, log_eps = packed
        log_prob = ops.dl_log_prob(y, raw, num_bins, log_eps).unsqueeze(-1)
    else:
        y_b, loc_b, ls_b = torch.broadcast_tensors(y, loc, log_scale)
        raw = torch.stack([loc_b, ls_b], dim=-1)
        log_prob = ops.dl_log_prob(y_b, raw, num_bins, _NO_CLAMP)
    return reduce(log_prob, reduce_dim) if reduce_dim else log_prob


def gaussian_ll(y, mu, sd, epsilon: float = 1e-6, reduce_dim: Optional[int] = -1):
    """Elementwise Gaussian log-likelihood — drop-in for blvm/utils/log_likelihoods.py:17-39.  The standard deviation is
    clamped at `epsilon`; the reference does that under no_grad, which detaches sd, so with epsilon != 0 no gradient
    reaches sd (kept).  `reduce_dim` falsy => no reduction."""
    if not isinstance(sd, torch.Tensor):
        sd = torch.as_tensor(float(sd), device=mu.device)
    log_prob = ops.gaussian_ll_elementwise(y, mu, sd, sd_floor=float(epsilon or 0.0))
    return reduce(log_prob, reduce_dim) if reduce_dim else log_prob


def gaussian_mixture_ll(y, logits, mu, sd, epsilon: float = 1e-6, reduce_dim: int = -1):
    """Gaussian-mixture log-likelihood — drop-in for blvm/utils/log_likelihoods.py:42-60.
    y (*, D); logits (*, K); mu, sd (*, D, K) -> (*).  Runs the DMoL tile kernel with the Gaussian component density."""
    if reduce_dim not in (-1, y.ndim - 1):
        raise NotImplementedError("blvm_b200.gaussian_mixture_ll reduces over the last (D) axis of y only")
    packed = getattr(sd, "_blvm_packed_gmm", None)
    if packed is not None and getattr(logits, "_blvm_packed_gmm", (None,))[0] is packed[0] and not epsilon \
            and y.shape[:-1] == packed[0].shape[:-1]:
        raw, K, D, beta, sd_add = packed
        return ops.gmm_log_prob(y, raw, K, D, True, beta, sd_add, 0.0)
    K, D = logits.size(-1), y.size(-1)
    batch = torch.broadcast_shapes(y.shape[:-1], logits.shape[:-1], mu.shape[:-2], sd.shape[:-2])
    raw = torch.cat([logits.expand(*batch, K), torch.cat([mu.expand(*batch, D, K), sd.expand(*batch, D, K)], dim=-1).flatten(-2)], dim=-1)
    return ops.gmm_log_prob(y.expand(*batch, D), raw, K, D, False, 1.0, 0.0, float(epsilon or 0.0))

"""Variational helpers with the reference's names (blvm/utils/variational.py): analytic Gaussian KL and free nats."""
import math
from typing import Optional, Tuple, Union

import torch

from . import ops

__all__ = ["kl_divergence_gaussian", "kl_divergence_gaussian_mc", "discount_free_nats"]


def kl_divergence_gaussian(mu_q: torch.Tensor, sd_q: torch.Tensor, mu_p: torch.Tensor, sd_p: torch.Tensor):
    """Elementwise analytic KL(q||p) between diagonal Gaussians given means and STANDARD DEVIATIONS — drop-in for
    blvm/utils/variational.py:67-70.  One kernel forward, one backward; evaluated as
    -log1p(rho-1) + ((rho-1)(rho+1) + z^2)/2 so that q ~ p does not cancel."""
    return ops.kl_gaussian(mu_q, sd_q, mu_p, sd_p)


def kl_divergence_gaussian_mc(mu_q, sd_q, mu_p, sd_p, z, epsilon: float = 0, reduce_dim: Optional[int] = None):
    """Elementwise Monte-Carlo KL log q(z) - log p(z) — drop-in for blvm/utils/variational.py:73-83 (bottom-up STCN)."""
    from .log_likelihoods import gaussian_ll
    return gaussian_ll(z, mu_q, sd_q, epsilon, reduce_dim) - gaussian_ll(z, mu_p, sd_p, epsilon, reduce_dim)


def discount_free_nats(kld: torch.Tensor, free_nats: float = None, shared_dims: Union[Tuple[int], int] = None):
    """max(kld, free_nats / prod(shape[shared_dims])) — blvm/utils/variational.py:86-122 (identity for None/0).

    A single elementwise op on an existing tensor: it stays a torch op here (its fused form lives inside
    `fused_elbo`, where max/mask/sum/gradient happen in the KL kernel)."""
    if free_nats is None or free_nats == 0:
        return kld
    if isinstance(shared_dims, int):
        shared_dims = (shared_dims,)
    if shared_dims is not None:
        min_kl_per_dim = free_nats / math.prod([kld.shape[d] for d in shared_dims])
    else:
        min_kl_per_dim = free_nats
    return torch.maximum(kld, torch.tensor(min_kl_per_dim, dtype=kld.dtype, device=kld.device))

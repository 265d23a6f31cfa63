"""Variational helpers with the reference's names (blvm/utils/variational.py): analytic Gaussian KL and free nats."""
import math
from typing import Optional, Tuple, Union

import torch

from . import ops

__all__ = ["kl_divergence_gaussian", "kl_divergence_gaussian_eager", "kl_divergence_gaussian_mc", "discount_free_nats", "LazyKL", "LazyResult"]


class LazyKL(torch.Tensor):
    """What the drop-in `kl_divergence_gaussian` returns: a tensor-shaped handle on the four Gaussian parameter tensors.

    The reference models compute `kld = kl_divergence_gaussian(enc_mu, enc_sd, prior_mu, prior_sd)` and hand it straight
    to `compute_elbo` (vrnn.py:336-342, srnn.py:268-272, clockwork_vae.py:307-318), which masks, discounts and sums it.
    Materialising the elementwise KL in between costs 20 B/element forward + 8 to reduce + 36 backward and three
    launches; the fused KL kernel does value, free nats, mask, row sums and all four gradients in one pass over
    32 B/element.  So the elementwise tensor is only produced if somebody actually reads it: shape / dtype / device are
    answered from metadata, every other use (`kld * mask`, `kld.sum()`, `torch.stack([kld])`, indexing ...) evaluates
    the KL kernel once (`ops.kl_gaussian`, differentiable) and carries on with the result.  The `compute_elbo`
    drop-ins (elbo.py) recognise an unread handle and route its inputs to `KLLevel(mu_q, sd_q, mu_p, sd_p)`.
    """

    @staticmethod
    def __new__(cls, mu_q, sd_q, mu_p, sd_p):
        shape = torch.broadcast_shapes(mu_q.shape, sd_q.shape, mu_p.shape, sd_p.shape)
        needs = torch.is_grad_enabled() and any(t.requires_grad for t in (mu_q, sd_q, mu_p, sd_p))
        r = torch.Tensor._make_wrapper_subclass(cls, shape, dtype=torch.float32, device=mu_q.device, requires_grad=needs)
        r._blvm_inputs = (mu_q, sd_q, mu_p, sd_p)
        r._blvm_value = None
        return r

    @property
    def kl_inputs(self):
        """(mu_q, sd_q, mu_p, sd_p) while the elementwise KL has not been read, else None."""
        return self._blvm_inputs if self._blvm_value is None else None

    def materialize(self) -> torch.Tensor:
        if self._blvm_value is None:
            self._blvm_value = ops.kl_gaussian(*self._blvm_inputs)
        return self._blvm_value

    def __repr__(self):
        state = "unread" if self._blvm_value is None else "materialised"
        return f"LazyKL(shape={tuple(self.shape)}, device={self.device}, {state})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _LAZY_METADATA:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        swap = lambda a: a.materialize() if isinstance(a, LazyKL) else a   # noqa: E731
        args = torch.utils._pytree.tree_map(swap, args)
        kwargs = torch.utils._pytree.tree_map(swap, kwargs)
        return func(*args, **kwargs)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # reached only by C++-side uses that bypass __torch_function__: same answer, evaluate and carry on
        swap = lambda a: a.materialize() if isinstance(a, LazyKL) else a   # noqa: E731
        return func(*torch.utils._pytree.tree_map(swap, args), **torch.utils._pytree.tree_map(swap, kwargs or {}))


class LazyResult(torch.Tensor):
    """A tensor-shaped promise: shape / dtype / device are known, the value is produced by `thunk()` the first time anything
    reads it (same mechanism as LazyKL).  Used for `sample()` / `mode()` on likelihood parameters whose Linear has not been
    evaluated (LinearDMoLParams): the reference models call both on every training step (vrnn.py:332-333) and store the
    results in their outputs, where a training loop never looks at them -- evaluating them would force the (B, T, 3K)
    parameter tensor into existence and undo the fused head."""

    @staticmethod
    def __new__(cls, shape, dtype, device, thunk):
        r = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device, requires_grad=False)
        r._blvm_thunk = thunk
        r._blvm_value = None
        return r

    def materialize(self) -> torch.Tensor:
        if self._blvm_value is None:
            self._blvm_value = self._blvm_thunk()
            self._blvm_thunk = None
        return self._blvm_value

    def __repr__(self):
        return f"LazyResult(shape={tuple(self.shape)}, device={self.device}, {'unread' if self._blvm_value is None else 'materialised'})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _LAZY_METADATA:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        swap = lambda a: a.materialize() if isinstance(a, (LazyResult, LazyKL)) else a   # noqa: E731
        return func(*torch.utils._pytree.tree_map(swap, args), **torch.utils._pytree.tree_map(swap, kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        swap = lambda a: a.materialize() if isinstance(a, (LazyResult, LazyKL)) else a   # noqa: E731
        return func(*torch.utils._pytree.tree_map(swap, args), **torch.utils._pytree.tree_map(swap, kwargs or {}))


_LAZY_DEVICE_TYPES = {"cuda"}   # (the CPU glue test adds "cpu" to exercise the handle with oracle-backed fake kernels)
_T = torch.Tensor
_LAZY_METADATA = {
    _T.shape.__get__, _T.dtype.__get__, _T.device.__get__, _T.ndim.__get__, _T.is_cuda.__get__, _T.requires_grad.__get__,
    _T.layout.__get__, _T.is_leaf.__get__, _T.grad_fn.__get__, _T.size, _T.dim, _T.numel, _T.ndimension, _T.nelement,
    _T.is_floating_point, _T.is_complex, _T.element_size, _T.__len__, _T.__repr__, _T.__str__, _T.__format__, _T.__hash__,
}


def kl_divergence_gaussian(mu_q: torch.Tensor, sd_q: torch.Tensor, mu_p: torch.Tensor, sd_p: torch.Tensor):
    """Elementwise analytic KL(q||p) between diagonal Gaussians given means and STANDARD DEVIATIONS — drop-in for
    blvm/utils/variational.py:67-70.  Evaluated as -log1p(rho-1) + ((rho-1)(rho+1) + z^2)/2 so that q ~ p does not
    cancel.  On CUDA the result is a `LazyKL`: it behaves like the (fp32) elementwise KL tensor — one
    kernel forward, one backward, when it is read — but `compute_elbo` consumes its inputs through the fused KL kernel
    without ever materialising it.  `kl_divergence_gaussian_eager` always evaluates."""
    if all(isinstance(t, torch.Tensor) and t.device.type in _LAZY_DEVICE_TYPES for t in (mu_q, sd_q, mu_p, sd_p)):
        return LazyKL(mu_q, sd_q, mu_p, sd_p)
    return ops.kl_gaussian(mu_q, sd_q, mu_p, sd_p)   # raises on CPU tensors: there is no CPU path


def kl_divergence_gaussian_eager(mu_q, sd_q, mu_p, sd_p):
    """The elementwise KL tensor itself (one kernel forward, one backward)."""
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

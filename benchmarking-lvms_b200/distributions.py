"""Likelihood modules with the reference's class names, constructor arguments, attributes, state-dict keys and method
contracts (blvm/modules/distributions.py:268-387) — the drop-in boundary of SURVEY.md §8b.

`forward` keeps the Linear (a cuBLAS GEMM, not on the hot path) and returns a parameter container that indexes like the
reference's `(logit_probs, locs, log_scales)` tuple but also carries the packed Linear output, so `log_prob` and the
fused ELBO read it once and clamp inside the kernel instead of materialising the clamped copy (distributions.py:386).
"""
from typing import Optional

import torch
import torch.nn as nn

from . import ops
import math

from .log_likelihoods import discretized_logistic_ll, discretized_logistic_mixture_ll, gaussian_mixture_ll

__all__ = ["ConditionalDistribution", "DiscretizedLogisticMixtureDense", "DiscretizedLogisticDense", "DMoLParams", "LinearDMoLParams",
           "DLParams", "DiagonalGaussianMixtureDense", "GMMParams"]


class ConditionalDistribution(nn.Module):
    """Same abstract surface as blvm/modules/distributions.py:28-52."""

    def reset_parameters(self):
        pass

    @staticmethod
    def get_distribution(*args, **kwargs):
        raise NotImplementedError()

    @staticmethod
    def sample(params):
        raise NotImplementedError()

    @staticmethod
    def rsample(params):
        raise NotImplementedError()

    @staticmethod
    def mode(params):
        raise NotImplementedError()

    def log_prob(self, x):
        raise NotImplementedError()


class DMoLParams:
    """What DiscretizedLogisticMixtureDense.forward returns: behaves like the reference's 3-tuple
    `(logit_probs (*, K), locs (*, D, K), log_scales (*, D, K) clamped at log_epsilon)` under indexing, iteration and
    len(), and keeps `raw (*, K(2D+1))`, the packed Linear output the kernels consume.  The clamped log-scales are only
    materialised if somebody asks for them (`params[2]`, e.g. `sample`)."""

    __slots__ = ("raw", "K", "D", "log_epsilon", "_cache")

    def __init__(self, raw: torch.Tensor, K: int, D: int, log_epsilon: float):
        self.raw, self.K, self.D, self.log_epsilon = raw, K, D, log_epsilon
        self._cache = {}

    def _tag(self, t):
        t._blvm_packed = (self.raw, self.K, self.D, self.log_epsilon)
        return t

    def __len__(self):
        return 3

    def __iter__(self):
        return iter((self[0], self[1], self[2]))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return tuple(self)[i]
        if i < 0:
            i += 3
        if i not in self._cache:
            raw, K, D = self.raw, self.K, self.D
            if i == 0:
                t = raw[..., :K]                                                    # distributions.py:383
            elif i in (1, 2):
                lls = raw[..., K:].view(*raw.shape[:-1], D, 2 * K)                  # :384
                t = lls[..., :K] if i == 1 else lls[..., K:].clamp(min=self.log_epsilon)  # :385-386
            else:
                raise IndexError(i)
            self._cache[i] = self._tag(t)
        return self._cache[i]

    def detach(self):
        return DMoLParams(self.raw.detach(), self.K, self.D, self.log_epsilon)


class LinearDMoLParams(DMoLParams):
    """DMoLParams whose packed Linear output has NOT been computed: it carries the Linear's input and parameters instead, so
    that `fused_elbo` can run the whole likelihood head -- `x W^T + b` on the tcgen05 tensor cores, the DMoL value and
    gradient, and the Linear's backward -- in one kernel without the (B, T, 3K) tensor ever reaching HBM
    (csrc/linear_dmol_kernel.cuh).  Anything else that wants the parameters (`params[i]`, `sample()`,
    `mode()`, `log_prob()`) reads `.raw`, which evaluates the Linear once (cuBLAS) and caches it: same values as the
    reference's `self.params(x)`, the fused path is then simply not taken."""

    __slots__ = ("x", "weight", "bias", "_raw")

    def __init__(self, x, weight, bias, K: int, D: int, log_epsilon: float):
        self.x, self.weight, self.bias = x, weight, bias
        self.K, self.D, self.log_epsilon = K, D, log_epsilon
        self._raw = None
        self._cache = {}

    @property
    def raw(self):
        if self._raw is None:
            # distributions.py:382 as autocast evaluates it: weight and bias cast to the activation dtype
            dt = self.x.dtype
            self._raw = torch.nn.functional.linear(self.x, self.weight.to(dt), None if self.bias is None else self.bias.to(dt))
        return self._raw

    @property
    def materialized(self) -> bool:
        return self._raw is not None

    def detach(self):
        return DMoLParams(self.raw.detach(), self.K, self.D, self.log_epsilon)


class DLParams:
    """Same idea for DiscretizedLogisticDense: indexes like `(mu (*, D), log_scale (*, D) clamped)`, carries raw (*, 2D)."""

    __slots__ = ("raw", "D", "log_epsilon", "_cache")

    def __init__(self, raw, D, log_epsilon):
        self.raw, self.D, self.log_epsilon = raw, D, log_epsilon
        self._cache = {}

    def __len__(self):
        return 2

    def __iter__(self):
        return iter((self[0], self[1]))

    def __getitem__(self, i):
        if i < 0:
            i += 2
        if i not in self._cache:
            mu, ls = self.raw.chunk(2, dim=-1)                                      # distributions.py:305
            t = mu if i == 0 else ls.clamp(min=self.log_epsilon)                    # :306
            if i not in (0, 1):
                raise IndexError(i)
            t._blvm_packed_dl = (self.raw, self.log_epsilon)
            self._cache[i] = t
        return self._cache[i]


def _rsample_logistic(mu, log_scale, eps: float = 1e-8):
    """mu + exp(log_scale) * logit(u), u ~ U(eps, 1-eps)  (blvm/utils/variational.py:282-294), then clamp to [-1, 1]
    (:296-306).  Sampling is not on the hot path: plain torch."""
    u = torch.empty_like(mu).uniform_(eps, 1 - eps)
    return (mu + torch.exp(log_scale) * (torch.log(u) - torch.log(1 - u))).clamp(-1, 1)


class DiscretizedLogisticDense(ConditionalDistribution):
    """Drop-in for blvm/modules/distributions.py:268-307 (state-dict keys `params.weight`, `params.bias`)."""

    def __init__(self, x_dim: int, y_dim: int, num_bins: int = 256, log_epsilon: float = -7.0):
        super().__init__()
        self.x_dim = x_dim
        self.y_dim = y_dim
        self.num_bins = num_bins
        self.log_epsilon = log_epsilon
        self.out_features = y_dim * 2
        self.params = nn.Linear(x_dim, self.out_features)
        self.reset_parameters()

    @staticmethod
    def rsample(params):
        return _rsample_logistic(params[0], params[1])

    @staticmethod
    @torch.no_grad()
    def sample(params):
        return _rsample_logistic(params[0], params[1])

    def mode(self, params):
        return params[0]

    def log_prob(self, y, params, reduce_dim: Optional[int] = None):
        """Inputs are assumed to be in [-1, 1]."""
        if isinstance(params, DLParams) and self.y_dim == 1 and y.shape == params.raw.shape[:-1] + (1,):
            log_prob = ops.dl_log_prob(y, params.raw, self.num_bins, self.log_epsilon).unsqueeze(-1)
            if reduce_dim:
                return log_prob.squeeze(reduce_dim) if log_prob.size(reduce_dim) == 1 else log_prob.sum(reduce_dim)
            return log_prob
        return discretized_logistic_ll(y, params[0], params[1], num_bins=self.num_bins, reduce_dim=reduce_dim)

    def forward(self, x):
        return DLParams(self.params(x), self.y_dim, self.log_epsilon)


class DiscretizedLogisticMixtureDense(ConditionalDistribution):
    """Drop-in for blvm/modules/distributions.py:310-387: `3 * num_mix` parameters per output channel
    (mixture logit, mean, log-scale); data assumed rescaled to `num_bins` discrete values in [-1, 1]."""

    def __init__(self, x_dim: int, y_dim: int, num_mix: int = 10, num_bins: int = 256, log_epsilon: float = -7.0,
                 fuse_linear: bool = False, lazy_samples: bool = False):
        super().__init__()
        self.x_dim = x_dim
        self.y_dim = y_dim
        self.num_mix = num_mix
        self.num_bins = num_bins
        self.log_epsilon = log_epsilon
        self.out_features = num_mix * (2 * y_dim + 1)
        self.params = nn.Linear(x_dim, self.out_features)
        # opt-in (not a reference argument): under AMP, hand `fused_elbo` the Linear's input instead of its output so that the
        # whole head runs as one tensor-core kernel (see LinearDMoLParams); everything else behaves as before
        self.fuse_linear = fuse_linear
        # opt-in (patch_blvm(lazy_samples=True)): sample() / mode() on CUDA parameters return promises (variational.LazyResult) that run
        # the fused sample + mode kernel the first time either of them is read -- the reference models call both on every training step
        # (vrnn.py:332-333) and store the results in outputs a training loop never looks at.  The random draw then happens at the first
        # read, and a lazily read mode() is detached (nothing in the reference differentiates it).
        self.lazy_samples = lazy_samples
        self.reset_parameters()

    @staticmethod
    def get_distribution(params):
        raise NotImplementedError("Discretized mixture of logistics does not have a Distribution object (yet)")

    def rsample(self, params):
        """Reparameterised sample: Gumbel-max over the mixture logits, gather the chosen component, sample its logistic
        (blvm/utils/variational.py:309-349 with the defaults eps=1e-5, hard argmax).  NOT on the path (no model, experiment
        or test of the reference calls it: they use sample(), which is the kernel above) and not a fallback for anything: a
        composition of differentiable torch ops kept for API completeness, on whatever device the parameters live."""
        logit_probs, locs, log_scales = params[0], params[1], params[2]
        u = torch.empty_like(logit_probs).uniform_(1e-5, 1.0 - 1e-5)
        choice = torch.argmax(logit_probs - torch.log(-torch.log(u)), dim=-1, keepdim=True)   # (*, 1)
        index = choice.expand(*choice.shape[:-1], locs.size(-2)).unsqueeze(-1)                 # (*, D, 1)
        loc = torch.gather(locs, index=index, dim=-1).squeeze(-1)
        log_scale = torch.gather(log_scales, index=index, dim=-1).squeeze(-1)
        return _rsample_logistic(loc, log_scale)

    def _fused_sample_mode(self, params):
        """One kernel reads the packed parameters once and produces both the sample and the mode; the result is cached
        on the parameter container because the models call sample() and mode() back to back (vrnn.py:332-333)."""
        cached = params._cache.get("sample_mode")
        if cached is None:
            cached = ops.dmol_sample_mode(params.raw, params.K, params.D, params.log_epsilon)
            params._cache["sample_mode"] = cached
        return cached

    @staticmethod
    def _as_packed(params) -> "DMoLParams":
        """The packed container the sample / mode kernel reads: what forward() returned, or a reference-style
        (logit_probs, locs, log_scales) tuple packed once (its log-scales are already clamped)."""
        if isinstance(params, DMoLParams):
            return params
        logit_probs, locs, log_scales = params[0], params[1], params[2]
        K, D = logit_probs.size(-1), locs.size(-2)
        raw = torch.cat([logit_probs, torch.cat([locs, log_scales], dim=-1).flatten(-2)], dim=-1)
        return DMoLParams(raw, K, D, -math.inf)

    def _deferred(self, params, which: int):
        """sample() / mode() on parameters whose Linear has not been evaluated (LinearDMoLParams): a promise of the result
        (variational.LazyResult), so that a training step that never looks at the reconstructions keeps the fused head."""
        from .variational import LazyResult
        x = params.x
        dtype = torch.float32 if (which == 0 or x.dtype == torch.float32) else x.dtype
        fn = self.sample if which == 0 else self.mode

        def thunk():
            params.raw                                     # evaluates the Linear (cuBLAS) once; the fused path is then not taken
            with torch.no_grad():
                return fn(params)
        return LazyResult((*x.shape[:-1], self.y_dim), dtype, x.device, thunk)

    def _lazy_sample_mode(self, params: "DMoLParams", which: int):
        """lazy_samples=True: a promise of sample() (which = 0) / mode() (1) on packed CUDA parameters; both come from ONE launch of the
        fused kernel, issued when the first of them is read (cached on the parameter container like the eager call)."""
        from .variational import LazyResult, _LAZY_DEVICE_TYPES
        raw = params.raw
        if raw.device.type not in _LAZY_DEVICE_TYPES:
            return self._fused_sample_mode(params)[which]      # raises for CPU tensors: no fallback
        dtype = torch.float32 if (which == 0 or raw.dtype == torch.float32) else raw.dtype

        def thunk():
            with torch.no_grad():
                out = self._fused_sample_mode(params)[which]
                return out if out.dtype == dtype else out.to(dtype)
        return LazyResult((*raw.shape[:-1], self.y_dim), dtype, raw.device, thunk)

    @torch.no_grad()
    def sample(self, params):
        """A sample of the mixture, clamped to [-1, 1] (distributions.py:359-361, variational.py:309-349): the fused
        sample + mode kernel.  CUDA only, like every kernel of this package (CPU tensors raise; no torch fallback)."""
        if isinstance(params, LinearDMoLParams) and not params.materialized:
            return self._deferred(params, 0)
        if self.lazy_samples:
            return self._lazy_sample_mode(self._as_packed(params), 0)
        return self._fused_sample_mode(self._as_packed(params))[0]

    def mode(self, params):
        """Mean of the most probable component (distributions.py:363-368), from the same kernel launch as sample();
        differentiable w.r.t. the chosen component's location like the reference's gather."""
        if isinstance(params, LinearDMoLParams) and not params.materialized:
            return self._deferred(params, 1)               # (read lazily, it is detached: nothing in the reference differentiates it)
        params = self._as_packed(params)
        if self.lazy_samples:
            return self._lazy_sample_mode(params, 1)
        _, mode, index = self._fused_sample_mode(params)
        if torch.is_grad_enabled() and params.raw.requires_grad:
            return ops.mode_with_grad(params.raw, mode, index, params.K, params.D)
        return mode if params.raw.dtype == torch.float32 else mode.to(params.raw.dtype)

    def log_prob(self, y, params, reduce_dim: int = -1):
        """Per-sample log-likelihood (*); inputs are assumed to be in [-1, 1] (distributions.py:370-379)."""
        if isinstance(params, DMoLParams) and y.shape == params.raw.shape[:-1] + (self.y_dim,) and reduce_dim in (-1, y.ndim - 1):
            return ops.dmol_log_prob(y, params.raw, self.num_mix, self.y_dim, self.num_bins, self.log_epsilon)
        return discretized_logistic_mixture_ll(y, params[0], params[1], params[2], num_bins=self.num_bins,
                                               reduce_dim=reduce_dim)

    def forward(self, x):
        if self.fuse_linear and x.is_cuda and x.dim() == 3 and self.y_dim == 1:
            # the dtype the Linear would compute in: the autocast dtype inside an autocast region, else x's own
            dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
            if dt in (torch.float16, torch.bfloat16) and ops.linear_dmol_supported(self.num_mix, self.x_dim):
                return LinearDMoLParams(x if x.dtype == dt else x.to(dt), self.params.weight, self.params.bias, self.num_mix,
                                        self.y_dim, self.log_epsilon)
        return DMoLParams(self.params(x), self.num_mix, self.y_dim, self.log_epsilon)


class GMMParams:
    """What DiagonalGaussianMixtureDense.forward returns: indexes like the reference's `(logit_probs (*, K),
    mu (*, D, K), sd (*, D, K))` with sd = softplus_beta(log_sd) + epsilon, and carries the packed Linear output so that
    log_prob / fused_elbo apply the activation (and its chain rule) inside the kernel."""

    __slots__ = ("raw", "K", "D", "beta", "sd_add", "_cache")

    def __init__(self, raw, K, D, beta, sd_add):
        self.raw, self.K, self.D, self.beta, self.sd_add = raw, K, D, beta, sd_add
        self._cache = {}

    def __len__(self):
        return 3

    def __iter__(self):
        return iter((self[0], self[1], self[2]))

    def __getitem__(self, i):
        if i < 0:
            i += 3
        if i not in self._cache:
            raw, K, D = self.raw, self.K, self.D
            if i == 0:
                t = raw[..., :K]                                                            # distributions.py:198
            elif i in (1, 2):
                mls = raw[..., K:].view(*raw.shape[:-1], D, 2 * K)                          # :199
                if i == 1:
                    t = mls[..., :K]
                else:                                                                       # :201 sd activation
                    t = torch.nn.functional.softplus(mls[..., K:], beta=self.beta) + self.sd_add
            else:
                raise IndexError(i)
            t._blvm_packed_gmm = (self.raw, self.K, self.D, self.beta, self.sd_add)
            self._cache[i] = t
        return self._cache[i]


class DiagonalGaussianMixtureDense(ConditionalDistribution):
    """Drop-in for blvm/modules/distributions.py:153-204 (`--likelihood GMM`): same constructor, attributes and
    state-dict keys; the sd activation Softplus(beta = ln2/initial_sd) + epsilon runs inside the likelihood kernel."""

    def __init__(self, x_dim, y_dim, num_mix: int, initial_sd: float = 1, epsilon: float = 1e-6):
        super().__init__()
        self.x_dim, self.y_dim, self.num_mix = x_dim, y_dim, num_mix
        self.initial_sd, self.epsilon = initial_sd, epsilon
        self.out_features = num_mix * (2 * y_dim + 1)
        self.params = nn.Linear(x_dim, self.out_features)
        # distributions.py:167-170: beta = ln2/initial_sd when epsilon > 0, ln2/(initial_sd - epsilon) otherwise
        self.softplus_beta = math.log(2) / (initial_sd if epsilon > 0 else (initial_sd - epsilon))
        self.reset_parameters()

    def rsample(self, params):
        """Gumbel-max component choice, then mu + sd * N(0, 1) (blvm/utils/variational.py:156-196); plain torch."""
        logits, mu, sd = params[0], params[1], params[2]
        u = torch.empty_like(logits).uniform_(1e-6, 1.0 - 1e-6)
        choice = torch.argmax(logits - torch.log(-torch.log(u)), dim=-1, keepdim=True)
        index = choice.expand(*choice.shape[:-1], mu.size(-2)).unsqueeze(-1)
        m = torch.gather(mu, index=index, dim=-1).squeeze(-1)
        s = torch.gather(sd, index=index, dim=-1).squeeze(-1)
        return torch.randn_like(m).mul(s).add(m)

    @torch.no_grad()
    def sample(self, params):
        return self.rsample(params)

    def mode(self, params):
        component = params[0].argmax(-1, keepdim=True).unsqueeze(-2)
        return torch.gather(params[1], index=component, dim=-1).squeeze(-1)

    def log_prob(self, y, params, reduce_dim: int = -1):
        if isinstance(params, GMMParams) and y.shape == params.raw.shape[:-1] + (self.y_dim,) and reduce_dim in (-1, y.ndim - 1):
            return ops.gmm_log_prob(y, params.raw, self.num_mix, self.y_dim, True, params.beta, params.sd_add, 0.0)
        return gaussian_mixture_ll(y, params[0], params[1], params[2], epsilon=0, reduce_dim=reduce_dim)

    def forward(self, x):
        return GMMParams(self.params(x), self.num_mix, self.y_dim, self.softplus_beta, self.epsilon if self.epsilon > 0 else 0.0)

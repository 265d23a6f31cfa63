"""Bits-per-dim / log-likelihood / KL running-mean arithmetic of blvm/evaluation/metrics.py:209-264,443-468 computed
from the sums the finalize kernel already produced: one device->host read for all of them instead of one
`.sum().tolist()` sync per Metric object (8-20 per step in the reference)."""
import math
from types import SimpleNamespace

import torch

__all__ = ["elbo_metrics", "bits_per_dim", "RunningMean", "patch_metric_syncs", "unpatch_metric_syncs", "tag_sum"]


def bits_per_dim(elbo: torch.Tensor, x_sl) -> float:
    """sum_b(-elbo_b / ln 2) / sum_b(x_sl_b) — BitsPerDimMetric(elbo, reduce_by=x_sl).value (metrics.py:456, :241-247)."""
    x_sl = torch.as_tensor(x_sl)
    return float((-elbo.detach().double() / math.log(2)).sum().item() / float(x_sl.sum()))


def elbo_metrics(result) -> SimpleNamespace:
    """One sync: read the 8 fp64 sums of a `fused_elbo` result and derive the per-step metric values the reference
    models log (vrnn.py:346-355): loss, elbo / rec / kl per utterance mean, kl in bits per timestep, bits per dim."""
    loss, s_logp, s_kl, s_klfn, s_elbo, s_len, bpd, _ = result.sums.tolist()
    B = result.elbo.numel()
    return SimpleNamespace(loss=loss, elbo=s_elbo / B, rec=s_logp / B, kl=s_kl / B, kl_fn=s_klfn / B,
                           kl_bpt=s_kl / math.log(2) / s_len, bpd=bpd, weight=s_len, batch=B)


class RunningMean:
    """Weighted running mean with the update rule of RunningMeanMetric.update (metrics.py:253-264)."""

    def __init__(self):
        self.value, self.weight = 0.0, 0.0

    def update(self, value: float, weight: float):
        d = self.weight + weight
        self.value = self.value * (self.weight / d) + value * (weight / d)
        self.weight = d
        return self.value


# ----------------------------------------------------------------------------------------------------------------------
# Sync-free Metric objects for the drop-in path
# ----------------------------------------------------------------------------------------------------------------------
# The reference models build 8-20 RunningMeanMetric objects per step (vrnn.py:346-355, stcn.py:236-253,
# clockwork_vae.py:99-127), and every constructor does `values.sum().tolist()` (metrics.py:241-244): one device->host
# synchronisation per metric.  Under `patch_blvm()` the constructor keeps the device-side sum instead (for the tensors
# `compute_elbo` returned that is an entry of the finalize kernel's 8 sums: no kernel at all) and queues it in the step's
# batch; the first metric whose value is actually read (`tracker.update(metrics)`, printing) fetches the WHOLE batch
# with one device->host copy.  Metric semantics (value / reduce_by, weight_by, the update rule) are untouched.


class _SyncBatch:
    """Device scalars queued since the last read; resolved together by ONE device->host copy."""

    def __init__(self):
        self.pending, self.values = [], None

    def add(self, t: torch.Tensor) -> int:
        self.pending.append(t)
        return len(self.pending) - 1

    def get(self, i: int) -> float:
        if self.values is None:
            self.values = torch.stack([t.reshape(()).to(torch.float64) for t in self.pending]).tolist()   # the step's one sync
            self.pending = None
        return self.values[i]


_LAZY_DEVICE_TYPES = {"cuda"}   # (the CPU glue test adds "cpu" to exercise the machinery without a GPU)
_batch = _SyncBatch()
_metric_patches = []
syncs_saved = 0   # device->host reads that were folded into a batch (diagnostic)


def tag_sum(t: torch.Tensor, sums: torch.Tensor, index: int) -> torch.Tensor:
    """Mark `t` (a per-utterance output of the fused op) as having its sum over utterances already available on the
    device as `sums[index]`: a lazily constructed metric then needs no reduction kernel for it."""
    t._blvm_sum = (sums, index)
    return t


class _Lazy:
    __slots__ = ("batch", "index")

    def __init__(self, t: torch.Tensor):
        global _batch, syncs_saved
        if _batch.values is not None:      # the previous step's batch has been read: start a new one
            _batch = _SyncBatch()
        src = getattr(t, "_blvm_sum", None)
        self.batch = _batch
        self.index = _batch.add(src[0][src[1]] if src is not None else t.sum())
        syncs_saved += 1

    def resolve(self) -> float:
        return self.batch.get(self.index)


def _number(x, default):
    """The constructor's `x.sum().tolist() if tensor else (x or default)` with CUDA tensors deferred."""
    if isinstance(x, torch.Tensor):
        return _Lazy(x.detach()) if x.device.type in _LAZY_DEVICE_TYPES else x.sum().tolist()
    return x or default


def _resolved(x):
    return x.resolve() if isinstance(x, _Lazy) else x


def _patch_metric_class(cls, base_init, value_attr, has_weight):
    """Lazy constructor + lazily settling properties for one Metric class whose constructor reduces tensors on the host
    (`values.sum().tolist()`): RunningMeanMetric (value in `running_mean`), EMAMetric (`ema`), LatestMeanMetric (`latest`)."""
    orig_init, orig_copy = cls.__init__, cls.copy

    def lazy_init(self, values, name, tags=None, reduce_by=None, weight_by=None, get_best=None, log_to_console=True,
                  log_to_framework=True):
        base_init(self, name=name, tags=tags, get_best=get_best, log_to_console=log_to_console, log_to_framework=log_to_framework)
        numel = values.numel() if isinstance(values, torch.Tensor) else 1              # metrics.py:240
        value = _number(values, values) if isinstance(values, torch.Tensor) else values
        reduce_by = _number(reduce_by, numel)                                           # :243
        weight_by = _number(weight_by, reduce_by) if has_weight else None               # :244
        d = self.__dict__
        if any(isinstance(v, _Lazy) for v in (value, reduce_by, weight_by)):
            d["_blvm_lazy"] = (value, reduce_by, weight_by)
        else:
            if has_weight:
                d["weight_by"] = weight_by
            d[value_attr] = value / reduce_by                                           # :246-247

    if not has_weight:   # LatestMeanMetric(values, name, tags, reduce_by, get_best, ...): no weight_by parameter
        def lazy_init(self, values, name, tags=None, reduce_by=None, get_best=None, log_to_console=True, log_to_framework=True,  # noqa: F811
                      _full=lazy_init):
            _full(self, values, name, tags=tags, reduce_by=reduce_by, get_best=get_best, log_to_console=log_to_console,
                  log_to_framework=log_to_framework)

    def _settle(self):
        lazy = self.__dict__.pop("_blvm_lazy", None)
        if lazy is not None:
            value, reduce_by, weight_by = (_resolved(v) for v in lazy)
            if has_weight:
                self.__dict__.setdefault("weight_by", weight_by)
            self.__dict__.setdefault(value_attr, value / reduce_by)

    def make_property(attr):
        def get(self):
            _settle(self)
            return self.__dict__[attr]

        def set_(self, v):
            _settle(self)
            self.__dict__[attr] = v
        return property(get, set_)

    def lazy_copy(self):
        _settle(self)
        return orig_copy(self)

    attrs = (value_attr, "weight_by") if has_weight else (value_attr,)
    cls.__init__ = lazy_init
    for attr in attrs:
        setattr(cls, attr, make_property(attr))
    cls.copy = lazy_copy
    _metric_patches.append((cls, orig_init, attrs))


def patch_metric_syncs(metrics_module):
    """Rebind the constructors of the reference's `blvm.evaluation.metrics` classes that the audio models build every step --
    RunningMeanMetric (and with it LossMetric / LLMetric / KLMetric / BitsPerDimMetric ...), EMAMetric (Clockwork-VAE,
    clockwork_vae.py:112) and LatestMeanMetric -- so that device values are read lazily: all metrics of a step, one sync."""
    if _metric_patches:
        return
    base_init = metrics_module.Metric.__init__
    _patch_metric_class(metrics_module.RunningMeanMetric, base_init, "running_mean", True)
    _patch_metric_class(metrics_module.EMAMetric, base_init, "ema", True)
    _patch_metric_class(metrics_module.LatestMeanMetric, base_init, "latest", False)


def unpatch_metric_syncs():
    while _metric_patches:
        cls, orig_init, attrs = _metric_patches.pop()
        cls.__init__ = orig_init
        if "copy" in cls.__dict__:
            del cls.copy
        for attr in attrs:
            if attr in cls.__dict__:
                delattr(cls, attr)

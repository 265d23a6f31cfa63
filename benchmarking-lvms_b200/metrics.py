"""Bits-per-dim / log-likelihood / KL running-mean arithmetic of blvm/evaluation/metrics.py:209-264,443-468 computed
from the sums the finalize kernel already produced: one device->host read for all of them instead of one
`.sum().tolist()` sync per Metric object (8-20 per step in the reference)."""
import math
from types import SimpleNamespace

import torch

__all__ = ["elbo_metrics", "bits_per_dim", "RunningMean"]


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

"""The masked ELBO reduction, fused: `fused_elbo` plus drop-in replacements for each reference model's
`compute_elbo` / `compute_loss` with identical signatures, return tuples, dtypes and quirks.

One step of the path = 1 DMoL kernel (value + gradient + masked row partials) + 1 KL kernel per latent level
(value + free nats + mask + gradient + row partials) + 1 finalize kernel.  `d loss / d log_prob[b, t] = -mask / sum(x_sl)`
is known before the kernels run (x_sl lives on the host), so parameter gradients are produced in the same pass and
`loss.backward()` only rescales them by the incoming grad_output (a no-op kernel when it is 1).
"""
import math
from types import SimpleNamespace
from typing import List, Optional, Sequence, Union

import torch

from . import amp, ops
from .distributions import DLParams, DMoLParams, GMMParams, LinearDMoLParams
from .log_likelihoods import _packed_source
from .metrics import tag_sum
from .operations import level_lengths, sequence_mask
from .variational import LazyKL

__all__ = ["KLLevel", "fused_elbo", "vrnn_compute_elbo", "srnn_compute_elbo", "cwvae_compute_elbo", "stcn_compute_loss",
           "wavenet_compute_loss", "pack_dmol_params"]


class KLLevel:
    """One latent layer's contribution to the ELBO.

    Either the four Gaussian parameter tensors `(mu_q, sd_q, mu_p, sd_p)`, each (B, Tz, Z) — KL, free nats, mask, sums
    and gradients then run in one kernel (with `z=` the level's KL is the Monte-Carlo estimate log q(z) - log p(z) of
    bottom-up STCN, same kernel) — or an already materialised elementwise `kld` (B, Tz, Z) as the reference's
    compute_elbo receives it.  `stride` = temporal stride of the layer relative to the waveform (valid steps =
    ceil(x_sl / stride)); alternatively explicit `lens` (B).  `free_nats` overrides the op-level budget for this level
    (Clockwork-VAE scales it per level, clockwork_vae.py:151).
    """

    def __init__(self, mu_q=None, sd_q=None, mu_p=None, sd_p=None, *, kld=None, z=None, stride: Optional[int] = None,
                 lens: Optional[torch.Tensor] = None, free_nats: Optional[float] = None):
        if kld is None and any(t is None for t in (mu_q, sd_q, mu_p, sd_p)):
            raise ValueError("KLLevel needs either (mu_q, sd_q, mu_p, sd_p) or kld=")
        if stride is None and lens is None:
            raise ValueError("KLLevel needs stride= or lens=")
        if kld is not None:
            self.tensors, self.kind = [kld], "kld"
        elif z is not None:   # Monte-Carlo KL log q(z) - log p(z) at the sample z (bottom-up STCN, variational.py:73-83)
            self.tensors, self.kind = [mu_q, sd_q, mu_p, sd_p, z], "mc"
        else:
            self.tensors, self.kind = [mu_q, sd_q, mu_p, sd_p], "inputs"
        self.stride, self.lens, self.free_nats = stride, lens, free_nats


def pack_dmol_params(parameters) -> "DMoLParams":
    """Accept what `likelihood(x)` returned: our DMoLParams (zero-copy) or a reference-style
    (logit_probs, locs, log_scales) tuple, which is packed into the (.., K(2D+1)) layout with one cat (the log-scales
    are then already clamped, so the kernel clamp is disabled)."""
    if isinstance(parameters, DMoLParams):
        return parameters
    logit_probs, locs, log_scales = parameters[0], parameters[1], parameters[2]
    src = _packed_source(logit_probs, locs, log_scales)   # all three must be views of the same Linear output
    if src is not None:
        return DMoLParams(*src)
    K, D = logit_probs.size(-1), locs.size(-2)
    raw = torch.cat([logit_probs, torch.cat([locs, log_scales], dim=-1).flatten(-2)], dim=-1)
    return DMoLParams(raw, K, D, -math.inf)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()   # AMP hands over fp16/bf16 Linear outputs; the kernels compute in fp32
    return t if t.is_contiguous() else t.contiguous()


def fused_elbo(
    y: torch.Tensor,
    parameters,
    x_sl: Union[torch.Tensor, Sequence[int]],
    kl_levels: Sequence[KLLevel] = (),
    beta: float = 1.0,
    free_nats: float = 0.0,
    *,
    num_bins: int,
    denom: Optional[float] = None,
    want_twise: bool = False,
    skip_padded: bool = False,
    x_sl_device: Optional[torch.Tensor] = None,
    grad_scaler=None,
    exchange=None,
    nansum: bool = False,
):
    """ELBO of a batch in one pass.

    Args:
        y: targets (B, T) or (B, T, D) in [-1, 1].
        parameters: `likelihood(x)` output (DMoLParams / DLParams / reference-style tuple) over (B, T), or None for a
            KL-only call.
        x_sl: valid samples per utterance (B), CPU int64 like in the reference (a CUDA tensor costs one sync unless
            `denom` is given).
        kl_levels: the latent layers (see KLLevel).
        beta, free_nats: as in `Model.forward(x, x_sl, beta=, free_nats=)`.
        num_bins: likelihood.num_bins.
        denom: normaliser of the loss; default sum(x_sl).  Data-parallel callers pass global_sum / world_size so that
            the mean of the per-rank losses (what DDP's gradient averaging implements) is the global loss.
        want_twise: also return the masked per-sample log-prob (B, T) (WaveNet returns it).
        skip_padded: do not read tiles that lie entirely in the padding (their outputs become exact zeros instead of
            `value * 0`; only differs from the reference if padded parameters are non-finite).
        x_sl_device: the same lengths already on the device (B) int64; with `denom` given the call then does no
            host<->device traffic at all and can be captured in a CUDA graph.
        grad_scaler: the `torch.amp.GradScaler` the caller will use for `scaler.scale(loss).backward()`.  Only matters for
            fp16 parameters: with a known scaler their gradient is written in the forward pass, pre-multiplied by the
            scaler's device-side scale (one pass instead of value + recompute, see amp.py).  Default: the single enabled
            scaler registered through `register_grad_scaler` / observed after `patch_blvm()`, if any.
        exchange: a `blvm_b200.SumsExchange`: the finalize kernel then also publishes this rank's sums to every rank
            over NVLink peer memory (`exchange.consume()` returns the global sums).
        nansum: WaveNet's reduction (wavenet.py:145): `loss = -nansum_b(log_prob_b) / sum(x_sl)`; utterances whose
            log-prob is NaN contribute nothing to the loss and receive a zero upstream gradient (their likelihood
            gradient rows are multiplied by 0 on the device, no host sync), the finite ones are unaffected.

    Returns a namespace with fp64 tensors: loss (), elbo, log_prob, kl, kl_fn (B,), kl_levels [(B,)],
    sums (8,) = [loss, sum log_prob, sum kl, sum kl_fn, sum elbo, sum x_sl, bits-per-dim, nansum-loss] and
    log_prob_twise (B, T) fp32 or None.  Only `loss` is differentiable (w.r.t. the likelihood parameters and the
    tensors of every KLLevel); the others are detached.
    """
    if isinstance(parameters, LinearDMoLParams):
        dev = parameters.x.device                       # (touching .raw would evaluate the Linear)
    elif parameters is not None:
        dev = parameters.raw.device if hasattr(parameters, "raw") else parameters[0].device
    else:
        dev = kl_levels[0].tensors[0].device
    x_sl_t = x_sl if isinstance(x_sl, torch.Tensor) else torch.as_tensor(x_sl)
    if x_sl_t.dtype != torch.int64:
        x_sl_t = x_sl_t.to(torch.int64)
    on_host = not x_sl_t.is_cuda
    total = float(denom) if denom is not None else float(x_sl_t.sum())
    B = x_sl_t.shape[0]

    # per-level valid lengths: computed where x_sl lives; host lengths travel with x_sl in ONE small H2D copy
    need = [lv for lv in kl_levels if lv.lens is None]
    if x_sl_device is not None:
        x_sl_dev = x_sl_device
        lens_of = {id(lv): level_lengths(x_sl_dev, int(lv.stride)) for lv in need}
    elif on_host:
        if need:
            packed = torch.stack([x_sl_t] + [level_lengths(x_sl_t, int(lv.stride)) for lv in need]).to(dev, non_blocking=True)
            x_sl_dev = packed[0]
            lens_of = {id(lv): packed[1 + i] for i, lv in enumerate(need)}
        else:
            x_sl_dev, lens_of = x_sl_t.to(dev, non_blocking=True), {}
    else:
        x_sl_dev = x_sl_t
        lens_of = {id(lv): level_lengths(x_sl_dev, int(lv.stride)) for lv in need}

    likelihood, raw, K, D, log_eps, gmm = "none", None, 1, 1, -7.0, (1.0, 0.0)
    head = None   # (x, weight, bias): the likelihood head runs fused (tensor-core Linear + DMoL + Linear backward in one kernel)
    if isinstance(parameters, LinearDMoLParams) and not parameters.materialized and not want_twise and not nansum and exchange is None \
            and parameters.x.dim() == 3 and parameters.x.shape[0] == B:
        p = parameters
        fp16 = p.x.dtype == torch.float16
        scaler = (grad_scaler if grad_scaler is not None else amp.active_grad_scaler(p.x.device)) if fp16 else None
        # fp16 gradients need the GradScaler's factor inside the kernel (amp.py); without a known scaler the head is evaluated unfused
        if not fp16 or not torch.is_grad_enabled() or amp.ensure_scale(scaler, p.x.device):
            head = (p.x, p.weight, p.bias, scaler)
            likelihood, K, D, log_eps = "linear_dmol", p.K, p.D, p.log_epsilon
            if y.numel() != B * p.x.shape[1]:
                raise ValueError(f"y {tuple(y.shape)} does not match the likelihood input {tuple(p.x.shape)}")
            y = _f32c(y)
    if parameters is not None and head is None:
        if isinstance(parameters, GMMParams):
            likelihood, raw, K, D, gmm = "gmm", parameters.raw, parameters.K, parameters.D, (parameters.beta, parameters.sd_add)
        elif isinstance(parameters, DLParams):
            likelihood, raw, log_eps = "dl", parameters.raw, parameters.log_epsilon
            if parameters.D != 1:
                raise NotImplementedError("fused_elbo supports DiscretizedLogisticDense with y_dim == 1")
        else:
            p = pack_dmol_params(parameters)
            likelihood, raw, K, D, log_eps = "dmol", p.raw, p.K, p.D, p.log_epsilon
        if raw.dim() != 3 or raw.shape[0] != B:
            raise ValueError(f"likelihood parameters must be (B, T, P) with B = len(x_sl) = {B}; got {tuple(raw.shape)}")
        raw = ops.param_tensor(raw, K, D) if likelihood == "dmol" else _f32c(raw)   # fp16/bf16 AMP outputs are read as is
        T = raw.shape[1]
        if y.numel() != B * T * D:
            raise ValueError(f"y {tuple(y.shape)} does not match parameters (B, T, D) = ({B}, {T}, {D})")
        y = _f32c(y)

    specs, flat = [], []
    for lv in kl_levels:
        ts = lv.tensors
        if len(ts) >= 4 and not all(t.shape == ts[0].shape for t in ts):
            ts = torch.broadcast_tensors(*ts)
        ts = [_f32c(t) for t in ts]
        if ts[0].dim() != 3 or ts[0].shape[0] != B:
            raise ValueError(f"KL tensors must be (B, Tz, Z) with B = {B}; got {tuple(ts[0].shape)}")
        lens = lv.lens if lv.lens is not None else lens_of[id(lv)]
        if lens.device != dev or lens.dtype != torch.int64:
            lens = lens.to(device=dev, dtype=torch.int64)
        fn = free_nats if lv.free_nats is None else lv.free_nats
        specs.append(ops.KLLevelSpec(lv.kind, float(fn or 0.0), lens, len(ts)))
        flat += ts

    need_grad = torch.is_grad_enabled() and ((raw is not None and raw.requires_grad) or any(t.requires_grad for t in flat)
                                             or (head is not None and any(t is not None and t.requires_grad for t in head[:3])))
    loss_scale = None
    if head is not None and need_grad and head[3] is not None:
        loss_scale = amp.scale_tensor_f64(head[3])
    if need_grad and likelihood == "dmol" and raw.dtype == torch.float16:
        scaler = grad_scaler if grad_scaler is not None else amp.active_grad_scaler(raw.device)
        if amp.ensure_scale(scaler, raw.device):
            loss_scale = amp.scale_tensor_f64(scaler)
    spec = ops.ELBOSpec(K=K, D=D, num_bins=int(num_bins), log_epsilon=float(log_eps), beta=float(beta), denom=total,
                        levels=specs, want_twise=want_twise, skip_padded=skip_padded, need_grad=need_grad,
                        likelihood=likelihood, exchange=exchange, gmm=gmm, loss_scale=loss_scale, nansum=bool(nansum))
    if head is not None:
        loss, sums, rows, twise = ops.fused_linear_elbo_apply(spec, y, x_sl_dev, head[0], head[1], head[2], flat)
    else:
        loss, sums, rows, twise = ops.fused_elbo_apply(spec, y, x_sl_dev, raw, flat)
    return SimpleNamespace(loss=loss, log_prob=rows[0], kl=rows[1], kl_fn=rows[2], elbo=rows[3],
                           kl_levels=[rows[4 + l] for l in range(len(specs))], sums=sums,
                           log_prob_twise=twise if want_twise else None, x_sl=x_sl_dev)


# ----------------------------------------------------------------------------------------------------------------------
# drop-in reducers (bound as methods by patch_blvm; `self` only needs the attributes the reference methods read)
# ----------------------------------------------------------------------------------------------------------------------
def _kl_level(kld, **kw) -> KLLevel:
    """A latent level from what the model hands to compute_elbo: the drop-in `kl_divergence_gaussian` returns a LazyKL;
    while nobody has read it, its four parameter tensors go to the fully fused KL kernel (one launch for all levels,
    32 B/element).  A real (or already read) elementwise KL tensor takes the materialised-KL path."""
    if isinstance(kld, LazyKL):
        inputs = kld.kl_inputs
        if inputs is not None:
            return KLLevel(*inputs, **kw)
        kld = kld.materialize()
    return KLLevel(kld=kld, **kw)


def _vrnn_like(self, y, parameters, kld_twise, x_sl, stride, beta, free_nats, return_fn_kl):
    r = fused_elbo(y, parameters, x_sl, [_kl_level(kld_twise, stride=stride)], beta, free_nats,
                   num_bins=self.likelihood.num_bins)
    # vrnn.py:266: dtype=float => float64; T = max(x_sl) from the host copy, the comparison against the lengths the fused
    # op already uploaded (no second host->device copy, no synchronisation)
    x_sl_host = x_sl if isinstance(x_sl, torch.Tensor) else torch.as_tensor(x_sl)
    if x_sl_host.is_cuda:
        seq_mask = sequence_mask(x_sl_host, dtype=torch.float64, device=y.device)
    else:
        seq_mask = (torch.arange(int(x_sl_host.max()), device=y.device).unsqueeze(0) < r.x_sl.unsqueeze(1)).to(torch.float64)
    kld = tag_sum(r.kl_fn, r.sums, 3) if return_fn_kl else tag_sum(r.kl, r.sums, 2)
    # the sums over utterances already exist on the device (finalize kernel): lazily built Metric objects use them
    return tag_sum(r.loss, r.sums, 0), tag_sum(r.elbo, r.sums, 4), tag_sum(r.log_prob, r.sums, 1), kld, seq_mask   # all float64 like the reference


def vrnn_compute_elbo(self, y, parameters, kld_twise, x_sl, stride: int, beta: float = 1, free_nats: float = 0):
    """Drop-in for VRNN.compute_elbo (blvm/models/vrnn.py:255-279): returns (loss, elbo, log_prob, kld, seq_mask), all
    float64.  Quirk kept: the returned `kld` is the free-nats-discounted KL (:276-279 overwrites it)."""
    return _vrnn_like(self, y, parameters, kld_twise, x_sl, stride, beta, free_nats, True)


def srnn_compute_elbo(self, y, parameters, kld_twise, x_sl, stride: int, beta: float = 1, free_nats: float = 0):
    """Drop-in for SRNN.compute_elbo (blvm/models/srnn.py:137-160): same, but returns the raw KL (:153,160)."""
    return _vrnn_like(self, y, parameters, kld_twise, x_sl, stride, beta, free_nats, False)


def _f32_outputs(r):
    """(loss, elbo, log_prob, kl, [kl per level]) in float32 (CW-VAE / STCN return types), each tagged with the device-side
    sum over utterances the finalize kernel already produced."""
    f = torch.float32
    return (tag_sum(r.loss.to(f), r.sums, 0), tag_sum(r.elbo.to(f), r.sums, 4), tag_sum(r.log_prob.to(f), r.sums, 1),
            tag_sum(r.kl.to(f), r.sums, 2), [k.to(f) for k in r.kl_levels])


def cwvae_compute_elbo(self, y, seq_mask, level_masks, x_sl, parameters, kld_layerwise: List[torch.Tensor],
                       beta: float = 1, free_nats: float = 0):
    """Drop-in for CWVAE.compute_elbo (blvm/models/clockwork_vae/clockwork_vae.py:132-161): float32 outputs,
    per-level free nats scaled by overall_strides[l] / overall_strides[0] (:151), KL summed over levels (:155).
    The (prefix) masks the caller built (:231-240) are turned back into lengths."""
    levels = []
    for l in range(self.num_levels):
        fn = free_nats * self.overall_strides[l] / self.overall_strides[0]
        levels.append(_kl_level(kld_layerwise[l], lens=level_masks[l].sum(1), free_nats=fn))
    r = fused_elbo(y, parameters, x_sl, levels, beta, free_nats, num_bins=self.likelihood.num_bins)
    return _f32_outputs(r)


def stcn_compute_loss(self, y, x_sl, parameters, mu_p, sd_p, mu_q, sd_q, z, free_nats: float, beta: float):
    """Drop-in for STCN.compute_loss (blvm/models/stcn/stcn.py:256-297).  Top-down (analytic KL, :286): every latent level
    goes through the fully fused KL kernel.  Bottom-up (:288): the level's KL is the Monte-Carlo estimate
    log q(z) - log p(z) (variational.py:73-83), evaluated by the same kernel from the four parameter tensors and the sample z
    (value, mask, free nats, sums and all five gradients in one pass).  mask -> free nats -> mask (:286-289) equals max(kl, fn/Z) on valid steps and 0 on padded ones, which is what
    the kernel computes (also for negative MC estimates)."""
    if self.top_down:
        levels = [KLLevel(mu_q[l], sd_q[l], mu_p[l], sd_p[l], stride=self.n_stack_frames) for l in range(self.n_latents)]
    else:
        levels = [KLLevel(mu_q[l], sd_q[l], mu_p[l], sd_p[l], z=z[l], stride=self.n_stack_frames) for l in range(self.n_latents)]
    r = fused_elbo(y, parameters, x_sl, levels, beta, free_nats, num_bins=self.likelihood_module.num_bins)
    return _f32_outputs(r)


def wavenet_compute_loss(self, y, x_sl, parameters):
    """Drop-in for WaveNet.compute_loss (blvm/models/wavenet/wavenet.py:128-146): (loss, log_prob (B,),
    log_prob_twise (B, T)); the loss uses nansum over utterances (:145)."""
    r = fused_elbo(y, parameters, x_sl, (), 1.0, 0.0, num_bins=self.likelihood.num_bins, want_twise=True, nansum=True)
    f = torch.float32
    return tag_sum(r.loss.to(f), r.sums, 0), tag_sum(r.log_prob.to(f), r.sums, 1), r.log_prob_twise

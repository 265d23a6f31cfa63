"""Tensor-level wrappers and autograd Functions over the C ABI (include/blvm_b200.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; all arithmetic of the path runs in the CUDA
kernels of libblvm_b200.so.  There is no CPU path: CPU tensors raise.
"""
import ctypes
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib
from ._lib import BLVM_FLAG_MASK_OUTPUT, BLVM_FLAG_SKIP_PADDED, check, lib

__all__ = [
    "dmol_log_prob", "dl_log_prob", "kl_gaussian", "KLLevelSpec", "ELBOSpec", "fused_elbo_apply", "quantize_indices", "dmol_sample_mode", "mode_with_grad", "gmm_log_prob", "gaussian_ll_elementwise",
    "check_input_range", "launch_count", "reset_launch_count",
]

_STRICT = os.environ.get("BLVM_B200_STRICT", "0") == "1"
_err_flags = {}
_launches = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def launch_count() -> int:
    return _launches


def reset_launch_count():
    global _launches
    _launches = 0


def _count(n=1):
    global _launches
    _launches += n


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                f"blvm_b200 kernels need CUDA tensors (got a tensor on {t.device}); there is no CPU fallback — "
                "use the reference implementation for CPU work.")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device_index: Optional[int] = None):
    """Raw cudaStream_t of torch's current stream (fast path avoids building a torch.cuda.Stream object)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device() if device_index is None else device_index)
    return torch.cuda.current_stream().cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager costs ~10 us)."""

    __slots__ = ("dev", "prev")

    def __init__(self, dev: torch.device):
        self.dev = dev

    def __enter__(self):
        idx = self.dev.index
        cur = torch.cuda.current_device()
        if idx is None or idx == cur:
            self.prev = None
        else:
            self.prev = cur
            torch.cuda.set_device(idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _sync_counter(device: torch.device) -> torch.Tensor:
    """Zero-initialised uint32 for blvm_elbo_finalize's last-block handshake, one per (device, stream)."""
    key = ("ctr", device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ctr = _err_flags.get(key)
    if ctr is None:
        ctr = torch.zeros(1, dtype=torch.int32, device=device)
        _err_flags[key] = ctr
    return ctr


def _err_flag(device: torch.device) -> torch.Tensor:
    key = device.index if device.index is not None else torch.cuda.current_device()
    flag = _err_flags.get(key)
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=device)
        _err_flags[key] = flag
    return flag


def check_input_range(device=None):
    """Raise AssertionError if any target passed to the DMoL/DL kernels since the last check was outside [-1, 1].

    The reference asserts this synchronously on every call (`assert torch.max(y) <= 1.0 and torch.min(y) >= -1.0`,
    blvm/utils/log_likelihoods.py:131,195), which costs two reductions and a device->host sync per step; the kernels
    record the violation in a device flag instead and this function reads it (one sync).  Set BLVM_B200_STRICT=1 to
    check after every call.
    """
    for key, flag in list(_err_flags.items()):
        if isinstance(key, tuple):
            continue
        if device is not None and torch.device(device).index not in (None, key):
            continue
        if int(flag.item()) != 0:
            flag.zero_()
            raise AssertionError("blvm_b200: targets outside [-1, 1] were passed to a discretized-logistic likelihood")


def _maybe_strict(device):
    if _STRICT:
        check_input_range(device)


_DTYPE_CODE = {torch.float32: _lib.BLVM_DTYPE_F32, torch.float16: _lib.BLVM_DTYPE_F16, torch.bfloat16: _lib.BLVM_DTYPE_BF16}


def param_tensor(raw: torch.Tensor, K: int, D: int) -> torch.Tensor:
    """Contiguous likelihood parameters in a dtype the kernels read directly: fp32 always, fp16/bf16 (the AMP Linear
    output) where a register kernel exists for (K, D); anything else is upcast to fp32 once."""
    if raw.dtype not in _DTYPE_CODE or (raw.dtype != torch.float32 and not lib.blvm_dmol_has_fast_path(K, D)):
        raw = raw.float()
    return raw if raw.is_contiguous() else raw.contiguous()


def _as_f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ----------------------------------------------------------------------------------------------------------------------
# raw kernel calls
# ----------------------------------------------------------------------------------------------------------------------
def _dmol_call(y, raw, x_sl_dev, gout, gscale, B, T, K, D, num_bins, log_eps, flags, lp, graw, partials, gscale_dev=None):
    _require_cuda(y, raw, x_sl_dev, gout, lp, graw, partials, gscale_dev)
    dt = _DTYPE_CODE[raw.dtype]
    with _on_device(raw.device):
        err = _err_flag(raw.device)
        if graw is None:
            rc = lib.blvm_dmol_fwd(_ptr(y), _ptr(raw), dt, _ptr(x_sl_dev), B, T, K, D, num_bins, log_eps, flags, _ptr(lp),
                                   _ptr(partials), _ptr(err), _stream())
            check(rc, "blvm_dmol_fwd")
        else:
            assert graw.dtype == raw.dtype
            rc = lib.blvm_dmol_fwd_grad(_ptr(y), _ptr(raw), dt, _ptr(x_sl_dev), _ptr(gout), gscale, _ptr(gscale_dev), B, T, K,
                                        D, num_bins, log_eps, flags, _ptr(lp), _ptr(graw), _ptr(partials), _ptr(err), _stream())
            check(rc, "blvm_dmol_fwd_grad")
    _count()


def _dl_call(y, raw, x_sl_dev, gout, gscale, B, T, num_bins, log_eps, flags, lp, graw, partials):
    _require_cuda(y, raw, x_sl_dev, gout, lp, graw, partials)
    with _on_device(raw.device):
        err = _err_flag(raw.device)
        rc = lib.blvm_dl_fwd_grad(_ptr(y), _ptr(raw), _ptr(x_sl_dev), _ptr(gout), gscale, B, T, num_bins, log_eps, flags,
                                  _ptr(lp), _ptr(graw), _ptr(partials), _ptr(err), _stream())
        check(rc, "blvm_dl_fwd_grad")
    _count()


# ----------------------------------------------------------------------------------------------------------------------
# per-sample log-prob with a generic autograd backward (recompute): the drop-in for `likelihood.log_prob(y, params)`
# ----------------------------------------------------------------------------------------------------------------------
class _DMoLLogProb(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, y, raw, K, D, num_bins, log_eps):
        y = _as_f32c(y)
        raw = param_tensor(raw, K, D)   # fp16/bf16 Linear outputs (AMP) are read as they are; arithmetic is fp32
        P = K * (2 * D + 1)
        assert raw.shape[-1] == P, f"raw last dim {raw.shape[-1]} != K(2D+1) = {P}"
        N = raw.numel() // P
        assert y.numel() == N * D, f"y has {y.numel()} elements, expected {N * D}"
        lp = torch.empty(raw.shape[:-1], dtype=torch.float32, device=raw.device)
        _dmol_call(y, raw, None, None, 0.0, 1, N, K, D, num_bins, log_eps, 0, lp, None, None)
        ctx.save_for_backward(y, raw)
        ctx.cfg = (K, D, num_bins, log_eps, N)
        _maybe_strict(raw.device)
        return lp

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        y, raw = ctx.saved_tensors
        K, D, num_bins, log_eps, N = ctx.cfg
        graw = torch.empty_like(raw)
        _dmol_call(y, raw, None, _as_f32c(g), 1.0, 1, N, K, D, num_bins, log_eps, 0, None, graw, None)
        return None, graw, None, None, None, None


def dmol_log_prob(y: torch.Tensor, raw: torch.Tensor, K: int, D: int, num_bins: int, log_epsilon: float) -> torch.Tensor:
    """log p(y) per element of the batch shape `raw.shape[:-1]` from the packed Linear output `raw (*, K(2D+1))`."""
    return _DMoLLogProb.apply(y, raw, K, D, num_bins, float(log_epsilon))


class _DLLogProb(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, y, raw, num_bins, log_eps):
        y = y.contiguous()
        raw = raw.contiguous()
        N = raw.numel() // 2
        assert y.numel() == N
        lp = torch.empty(raw.shape[:-1], dtype=torch.float32, device=raw.device)
        _dl_call(y, raw, None, None, 0.0, 1, N, num_bins, log_eps, 0, lp, None, None)
        ctx.save_for_backward(y, raw)
        ctx.cfg = (num_bins, log_eps, N)
        _maybe_strict(raw.device)
        return lp

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        y, raw = ctx.saved_tensors
        num_bins, log_eps, N = ctx.cfg
        graw = torch.empty_like(raw)
        _dl_call(y, raw, None, _as_f32c(g), 1.0, 1, N, num_bins, log_eps, 0, None, graw, None)
        return None, graw, None, None


def dl_log_prob(y: torch.Tensor, raw: torch.Tensor, num_bins: int, log_epsilon: float) -> torch.Tensor:
    """Elementwise discretized-logistic log-prob from packed `raw (*, 2) = [mu | log_scale]`, y (*)."""
    return _DLLogProb.apply(y, raw, num_bins, float(log_epsilon))


# ----------------------------------------------------------------------------------------------------------------------
# sibling likelihoods: Gaussian mixture (packed layout) and elementwise Gaussian
# ----------------------------------------------------------------------------------------------------------------------
def _gmm_call(y, raw, x_sl_dev, gout, gscale, B, T, K, D, from_raw, beta, sd_add, sd_floor, flags, lp, graw, partials,
              gscale_dev=None):
    _require_cuda(y, raw, x_sl_dev, gout, lp, graw, partials, gscale_dev)
    with _on_device(raw.device):
        rc = lib.blvm_gmm_fwd_grad(_ptr(y), _ptr(raw), _ptr(x_sl_dev), _ptr(gout), gscale, _ptr(gscale_dev), B, T, K, D,
                                   1 if from_raw else 0, float(beta), float(sd_add), float(sd_floor), flags, _ptr(lp),
                                   _ptr(graw), _ptr(partials), _stream())
        check(rc, "blvm_gmm_fwd_grad")
    _count()


class _GMMLogProb(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, y, raw, K, D, from_raw, beta, sd_add, sd_floor):
        y = y.contiguous()
        raw = raw.contiguous()
        P = K * (2 * D + 1)
        assert raw.shape[-1] == P, f"raw last dim {raw.shape[-1]} != K(2D+1) = {P}"
        N = raw.numel() // P
        assert y.numel() == N * D
        lp = torch.empty(raw.shape[:-1], dtype=torch.float32, device=raw.device)
        _gmm_call(y, raw, None, None, 0.0, 1, N, K, D, from_raw, beta, sd_add, sd_floor, 0, lp, None, None)
        ctx.save_for_backward(y, raw)
        ctx.cfg = (K, D, from_raw, beta, sd_add, sd_floor, N)
        return lp

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        y, raw = ctx.saved_tensors
        K, D, from_raw, beta, sd_add, sd_floor, N = ctx.cfg
        graw = torch.empty_like(raw)
        _gmm_call(y, raw, None, _as_f32c(g), 1.0, 1, N, K, D, from_raw, beta, sd_add, sd_floor, 0, None, graw, None)
        return None, graw, None, None, None, None, None, None


def gmm_log_prob(y, raw, K, D, from_raw=True, beta=1.0, sd_add=0.0, sd_floor=0.0):
    """Gaussian-mixture log p(y) per element of `raw.shape[:-1]` from packed raw (*, K(2D+1)) = [logits | mu | s]."""
    return _GMMLogProb.apply(y, raw, K, D, bool(from_raw), float(beta), float(sd_add), float(sd_floor))


class _GaussianLL(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, y, mu, sd, sd_floor):
        y, mu, sd = (t.contiguous() for t in (y, mu, sd))
        _require_cuda(y, mu, sd)
        lp = torch.empty_like(mu)
        with _on_device(mu.device):
            check(lib.blvm_gaussian_ll(_ptr(y), _ptr(mu), _ptr(sd), None, mu.numel(), sd_floor, _ptr(lp), None, None, _stream()),
                  "blvm_gaussian_ll")
        _count()
        ctx.save_for_backward(y, mu, sd)
        ctx.sd_floor = sd_floor
        return lp

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        y, mu, sd = ctx.saved_tensors
        g = _as_f32c(g)
        g_mu, g_sd = torch.empty_like(mu), torch.empty_like(mu)
        with _on_device(mu.device):
            check(lib.blvm_gaussian_ll(_ptr(y), _ptr(mu), _ptr(sd), _ptr(g), mu.numel(), ctx.sd_floor, None, _ptr(g_mu), _ptr(g_sd),
                                       _stream()), "blvm_gaussian_ll (backward)")
        _count()
        return -g_mu, g_mu, (g_sd if ctx.sd_floor == 0 else None), None   # d/dy = -d/dmu


def gaussian_ll_elementwise(y, mu, sd, sd_floor: float = 0.0):
    y, mu, sd = torch.broadcast_tensors(y, mu, sd)
    return _GaussianLL.apply(y, mu, sd, float(sd_floor))


# ----------------------------------------------------------------------------------------------------------------------
# elementwise Gaussian KL
# ----------------------------------------------------------------------------------------------------------------------
class _KLGaussian(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, mu_q, sd_q, mu_p, sd_p):
        ins = [t.contiguous() for t in (mu_q, sd_q, mu_p, sd_p)]
        _require_cuda(*ins)
        kl = torch.empty_like(ins[0])
        with _on_device(kl.device):
            check(lib.blvm_kl_gaussian_fwd(*[_ptr(t) for t in ins], kl.numel(), _ptr(kl), _stream()), "blvm_kl_gaussian_fwd")
        _count()
        ctx.save_for_backward(*ins)
        return kl

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        ins = ctx.saved_tensors
        g = _as_f32c(g)
        outs = [torch.empty_like(ins[0]) for _ in range(4)]
        with _on_device(g.device):
            check(lib.blvm_kl_gaussian_bwd(*[_ptr(t) for t in ins], _ptr(g), g.numel(), *[_ptr(t) for t in outs], _stream()),
                  "blvm_kl_gaussian_bwd")
        _count()
        return tuple(outs)


def kl_gaussian(mu_q, sd_q, mu_p, sd_p):
    mu_q, sd_q, mu_p, sd_p = torch.broadcast_tensors(mu_q, sd_q, mu_p, sd_p)
    return _KLGaussian.apply(mu_q, sd_q, mu_p, sd_p)


# ----------------------------------------------------------------------------------------------------------------------
# the fused ELBO op: DMoL value+grad, KL value+grad per level, finalize — gradients are produced in the forward pass
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class KLLevelSpec:
    kind: str                      # "inputs" (mu_q, sd_q, mu_p, sd_p: fully fused), "mc" (+ z: Monte-Carlo KL, fully fused) or "kld" (elementwise KL given)
    free_nats: float               # budget per latent step for this level (already scaled, clockwork_vae.py:151)
    lens: Optional[torch.Tensor]   # (B) int64 on device: valid latent steps, None = all
    n_tensors: int = 4


@dataclass
class ELBOSpec:
    K: int
    D: int
    num_bins: int
    log_epsilon: float
    beta: float
    denom: float                   # sum(x_sl) (global / world_size under data parallelism)
    levels: List[KLLevelSpec] = field(default_factory=list)
    want_twise: bool = False
    skip_padded: bool = False
    need_grad: bool = True
    likelihood: str = "dmol"       # "dmol" | "dl" | "gmm" | "none"
    gmm: tuple = (1.0, 0.0)        # (softplus beta, sd epsilon) of the Gaussian-mixture sd activation
    nansum: bool = False           # WaveNet.compute_loss: loss = -nansum(logp)/sum(x_sl), NaN utterances get a zero gradient
    exchange: object = None        # distributed.SumsExchange: publish the sums to all ranks from the finalize kernel
    loss_scale: Optional[torch.Tensor] = None   # fp64 device scalar (a GradScaler's scale): fp16 gradients in ONE pass, pre-scaled


_unit_grads = {}
_LIK_CODE = {"none": _lib.BLVM_LIK_NONE, "dmol": _lib.BLVM_LIK_DMOL, "dl": _lib.BLVM_LIK_DL, "gmm": _lib.BLVM_LIK_GMM}
_chunk_cache = {}


def _logp_chunks(likelihood: str, T: int, K: int, D: int) -> int:
    """Partial sums per utterance of the likelihood kernel (a function of the tile size the library picks for K)."""
    key = (likelihood, T, K, D)
    n = _chunk_cache.get(key)
    if n is None:
        n = int({"dmol": lambda: lib.blvm_dmol_chunks(T, K, D), "dl": lambda: lib.blvm_dl_chunks(T),
                 "gmm": lambda: lib.blvm_gmm_chunks(T, K, D)}[likelihood]())
        if len(_chunk_cache) < 4096:
            _chunk_cache[key] = n
    return n


def _unit_grad(device: torch.device) -> torch.Tensor:
    """A persistent fp64 scalar 1.0 per device: the implicit gradient of `loss.backward()`."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    t = _unit_grads.get(key)
    if t is None:
        t = torch.ones((), dtype=torch.float64, device=device)
        _unit_grads[key] = t
    return t


class FusedLoss(torch.Tensor):
    """The scalar loss returned by the fused ELBO op.  An ordinary tensor in every respect except that a plain
    `loss.backward()` hands autograd a persistent device-resident 1.0 as the root gradient (instead of letting it
    allocate and fill a fresh `ones_like`), which `_FusedELBO.backward` recognises: the gradients written by the forward
    pass are already d loss / d input, so neither the fill kernel nor the rescale launch runs.  Any other use
    (`scaler.scale(loss).backward()`, `(loss + reg).backward()`, an explicit `gradient=`) goes through the generic
    device-side rescale.  Results of arithmetic on it are plain tensors."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    def backward(self, gradient=None, retain_graph=None, create_graph=False, inputs=None):
        if gradient is None and self.is_cuda and self.dtype == torch.float64 and self.dim() == 0:
            gradient = _unit_grad(self.device)
        return torch.Tensor.backward(self, gradient, retain_graph, create_graph, inputs)


class _FusedELBO(torch.autograd.Function):
    """inputs: spec, y (B,T[,D]), x_sl_dev (B) int64, raw (B,T,P), then the KL tensors of every level, flattened.
    outputs: loss () fp64 [differentiable], scalars (8) fp64, rows (4+L, B) fp64, log_prob_twise (B,T) or empty.

    Host-side cost matters here (a step is ~200 us of GPU time): one fp64 workspace allocation holds the outputs and
    all per-tile partial sums, and the step is ONE call into the library (`blvm_elbo_step`: likelihood, the KL of all
    levels in one launch, finalize = 3 launches); backward launches nothing on a plain `loss.backward()`."""

    @staticmethod
    def forward(ctx, spec: ELBOSpec, y, x_sl_dev, raw, *kl_tensors):
        dev = raw.device if raw is not None else kl_tensors[0].device
        _require_cuda(y, x_sl_dev, raw, *kl_tensors)
        B = int(x_sl_dev.shape[0])
        L = len(spec.levels)
        assert L <= _lib.BLVM_MAX_KL_LEVELS, f"at most {_lib.BLVM_MAX_KL_LEVELS} KL levels"
        has_lik = spec.likelihood != "none"
        T = raw.shape[1] if has_lik else 0
        st = _lib.ElboStepStruct()
        st.likelihood = _LIK_CODE[spec.likelihood]
        st.K, st.D, st.num_bins, st.log_epsilon = spec.K, spec.D, spec.num_bins, spec.log_epsilon
        st.B, st.T, st.n_levels = B, T, L
        st.beta, st.denom = spec.beta, spec.denom
        st.x_sl = x_sl_dev.data_ptr()
        n_ws = 8 + (4 + L) * B
        grads: List[Optional[torch.Tensor]] = []
        twise = torch.empty(B, T, dtype=torch.float32, device=dev) if (has_lik and spec.want_twise) else torch.empty(0, device=dev)
        prescaled = deferred = False
        if has_lik:
            n_ws += B * _logp_chunks(spec.likelihood, T, spec.K, spec.D)
            # fp16 parameters (AMP with a GradScaler): gradients of magnitude ~1/sum(x_sl) would underflow in fp16 before
            # the loss scale is applied.  With a known scaler (amp.py) the kernel multiplies its device-side scale in and
            # writes the fp16 gradient pre-scaled in this pass; otherwise the gradient is produced in backward (second
            # launch, with the upstream grad_output read from the device).  fp32 / bf16 parameters get it in this pass.
            prescaled = spec.need_grad and raw.dtype == torch.float16 and spec.loss_scale is not None
            deferred = spec.need_grad and raw.dtype == torch.float16 and not prescaled
            graw = torch.empty_like(raw) if (spec.need_grad and not deferred) else None
            st.raw_dtype = _DTYPE_CODE[raw.dtype]
            st.flags = (BLVM_FLAG_MASK_OUTPUT | (BLVM_FLAG_SKIP_PADDED if spec.skip_padded else 0)
                        | (_lib.BLVM_FLAG_NANSUM_LOSS if spec.nansum else 0))
            st.y, st.raw = y.data_ptr(), raw.data_ptr()
            st.lp_twise = twise.data_ptr() if spec.want_twise else None
            st.graw = _ptr(graw)
            st.loss_scale = spec.loss_scale.data_ptr() if prescaled else None
            st.gmm_softplus_beta, st.gmm_sd_add = spec.gmm
            st.err_flag = _err_flag(dev).data_ptr()
            if deferred:
                ctx.save_for_backward(y, raw, x_sl_dev)
                ctx.deferred = (B, T, spec.K, spec.D, spec.num_bins, spec.log_epsilon, st.flags & 3, -1.0 / spec.denom)
            grads.append(graw)
        else:
            grads.append(None)

        i = 0
        for li, lv in enumerate(spec.levels):
            ts = kl_tensors[i:i + lv.n_tensors]
            i += lv.n_tensors
            Bz, Tz, Z = ts[0].shape
            assert Bz == B, f"KL level batch {Bz} != {B}"
            n_ws += 2 * B * (-(-(Tz * Z) // _lib.BLVM_KL_TILE))
            d = st.levels[li]
            d.lens, d.Tz, d.Z, d.free_nats = _ptr(lv.lens), Tz, Z, lv.free_nats
            if lv.kind == "kld":
                gk = torch.empty_like(ts[0]) if spec.need_grad else None
                d.kl, d.g_kl = ts[0].data_ptr(), _ptr(gk)
                grads.append(gk)
                continue
            n_g = 5 if lv.kind == "mc" else 4
            g = [torch.empty_like(ts[0]) for _ in range(n_g)] if spec.need_grad else [None] * n_g
            d.mu_q, d.sd_q, d.mu_p, d.sd_p = ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), ts[3].data_ptr()
            d.g_mu_q, d.g_sd_q, d.g_mu_p, d.g_sd_p = _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), _ptr(g[3])
            if lv.kind == "mc":
                d.z, d.g_z = ts[4].data_ptr(), _ptr(g[4])
            grads += g

        ws = torch.empty(n_ws, dtype=torch.float64, device=dev)   # outputs + partial sums, one allocation
        scalars = ws[:8]
        rows = ws[8:8 + (4 + L) * B].view(4 + L, B)
        st.workspace = ws.data_ptr()
        ex = spec.exchange
        if ex is not None:   # finalize + all-gather of the sums over NVLink peer memory, one kernel
            st.rank, st.world = ex.rank, ex.world
            st.peer_bases_host = (ctypes.c_void_p * ex.world)(*ex.peer_ptrs)
            st.exchange_counters, st.prev_global_sums, st.exchange_err = ex.counters.data_ptr(), ex.global_sums.data_ptr(), ex.err.data_ptr()
        with _on_device(dev):
            st.sync_counter = _sync_counter(dev).data_ptr()
            # ONE call: likelihood kernel, the KL of all levels (one launch, overlapping the likelihood grid's tail),
            # finalize (a programmatic dependent) and, for WaveNet's nansum, the NaN-row gate
            check(lib.blvm_elbo_step(ctypes.byref(st), _stream(dev.index)), "blvm_elbo_step")
        _count(lib.blvm_last_step_launches())   # as counted by the library: likelihood, KL of all levels, finalize (+ NaN-row gate)

        loss = scalars[:1].view(())
        ctx.set_materialize_grads(False)   # no zero-filled grads for the detached outputs
        if not deferred:
            ctx.deferred = None
        ctx.grads = grads
        ctx.prescale = spec.loss_scale if prescaled else None
        ctx.consumed = False
        ctx.mark_non_differentiable(scalars, rows, twise)
        _maybe_strict(dev)
        return loss, scalars, rows, twise

    @staticmethod
    def backward(ctx, g_loss, *unused):
        if ctx.consumed:
            raise RuntimeError("blvm_b200 fused ELBO: backward called twice (gradients are produced in the forward pass "
                               "and scaled in place; call the op again instead of retain_graph=True)")
        ctx.consumed = True
        if g_loss is None:
            return (None, None, None) + tuple(None for _ in ctx.grads)
        g = g_loss if (g_loss.dtype == torch.float64 and g_loss.is_contiguous()) else g_loss.to(torch.float64).contiguous()
        grads = list(ctx.grads)
        unit = _unit_grads.get(g.device.index)
        is_unit = unit is not None and g.data_ptr() == unit.data_ptr()   # FusedLoss.backward(): the root gradient is the constant 1

        def rescale(bufs, factor):
            n = len(bufs)
            with _on_device(g.device):
                rc = lib.blvm_scale_inplace_multi((ctypes.c_void_p * n)(*[b.data_ptr() for b in bufs]),
                                                  (ctypes.c_int64 * n)(*[b.numel() for b in bufs]),
                                                  (ctypes.c_int * n)(*[_DTYPE_CODE[b.dtype] for b in bufs]), n, factor.data_ptr(),
                                                  _stream())
                check(rc, "blvm_scale_inplace_multi")
            _count()

        if ctx.prescale is not None:
            # every gradient of the step (likelihood and KL levels) already carries the loss scale S: multiply by grad_output / S, which
            # is exactly 1 for scaler.scale(loss).backward() -- the one launch then exits at once
            bufs = [b for b in grads if b is not None]
            if bufs:
                rescale(bufs, g / ctx.prescale)
        else:
            bufs = [b for b in grads if b is not None]
            if bufs and not is_unit:
                rescale(bufs, g)
        if ctx.deferred is not None:   # fp16 parameters: value+gradient launch now, scaled by the device-side grad_output
            y, raw, x_sl_dev = ctx.saved_tensors
            B, T, K, D, num_bins, log_eps, flags, gscale = ctx.deferred
            graw = torch.empty_like(raw)
            _dmol_call(y, raw, x_sl_dev, None, gscale, B, T, K, D, num_bins, log_eps, flags, None, graw, None, gscale_dev=g)
            grads[0] = graw
        ctx.grads = None
        return (None, None, None) + tuple(grads)


def linear_dmol_supported(K: int, x_dim: int) -> bool:
    """Whether the fused likelihood head (Linear + DMoL + Linear backward in one tcgen05 kernel) is instantiated for this shape."""
    return bool(lib.blvm_linear_dmol_padded_dim(int(K), int(x_dim)))


_dw_partials = {}


def _dw_partial_buffer(device: torch.device, dp: int):
    """Per-(device, stream, padded x_dim) scratch for the per-CTA dW / db partials of the fused head (stream-ordered reuse)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream(), dp)
    buf = _dw_partials.get(key)
    if buf is None:
        buf = torch.empty(int(lib.blvm_linear_dmol_max_ctas()) * 32 * dp, dtype=torch.float32, device=device)
        _dw_partials[key] = buf
    return buf


class _FusedLinearELBO(torch.autograd.Function):
    """The fused ELBO op with the likelihood HEAD fused in: inputs spec, y (B,T), x_sl_dev, x (B,T,Din) fp16/bf16, weight (3K,Din),
    bias (3K) (the fp32 parameters of `likelihood.params`), then the KL tensors.  One tensor-core kernel computes x W^T + b, the
    DMoL value and gradient and dx / dW / db (csrc/linear_dmol_kernel.cuh); the KL of all levels and the finalize kernel follow
    as in _FusedELBO.  Outputs and backward contract as _FusedELBO (gradients are produced in the forward pass)."""

    @staticmethod
    def forward(ctx, spec: ELBOSpec, y, x_sl_dev, x, weight, bias, *kl_tensors):
        dev = x.device
        _require_cuda(y, x_sl_dev, x, weight, bias, *kl_tensors)
        B, T, Din = x.shape
        L = len(spec.levels)
        K = spec.K
        x = x if x.is_contiguous() else x.contiguous()
        w16 = weight.detach().to(x.dtype).contiguous()        # what autocast hands the Linear (distributions.py:382 under AMP)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        dp = int(lib.blvm_linear_dmol_padded_dim(K, Din))
        assert dp > 0, "unsupported shape for the fused head"
        chunks = (T + 127) // 128
        shapes, i = [], 0
        for lv in spec.levels:
            Bz, Tz, Z = kl_tensors[i].shape
            assert Bz == B
            shapes.append((Tz, Z, -(-(Tz * Z) // _lib.BLVM_KL_TILE)))
            i += lv.n_tensors
        n_out = 8 + (4 + L) * B
        ws = torch.empty(n_out + B * chunks + sum(2 * B * c for _, _, c in shapes), dtype=torch.float64, device=dev)
        base = ws.data_ptr()
        scalars, rows = ws[:8], ws[8:n_out].view(4 + L, B)
        off = n_out
        need = spec.need_grad
        dx = torch.empty_like(x) if need else None
        dwp = _dw_partial_buffer(dev, dp) if need else None
        twise = torch.empty(B, T, dtype=torch.float32, device=dev) if spec.want_twise else torch.empty(0, device=dev)
        grads: List[Optional[torch.Tensor]] = []
        with _on_device(dev):
            stream = _stream(dev.index)
            logp_ptr = base + 8 * off
            off += B * chunks
            used = ctypes.c_int64(0)
            flags = BLVM_FLAG_MASK_OUTPUT
            rc = lib.blvm_linear_dmol_fwd_grad(y.data_ptr(), x.data_ptr(), w16.data_ptr(), _ptr(b32), _DTYPE_CODE[x.dtype], x_sl_dev.data_ptr(),
                                               -1.0 / spec.denom, spec.loss_scale.data_ptr() if spec.loss_scale is not None else None,
                                               B, T, Din, K, spec.num_bins, spec.log_epsilon, flags, twise.data_ptr() if spec.want_twise else None,
                                               _ptr(dx), _ptr(dwp), int(lib.blvm_linear_dmol_max_ctas()) if need else 0, logp_ptr,
                                               _err_flag(dev).data_ptr(), None, ctypes.byref(used), stream)
            check(rc, "blvm_linear_dmol_fwd_grad")
            _count()
            dW = db = None
            if need:
                dW = torch.empty(3 * K, Din, dtype=torch.float32, device=dev)
                db = torch.empty(3 * K, dtype=torch.float32, device=dev)
                check(lib.blvm_linear_dmol_reduce_dw(dwp.data_ptr(), used.value, Din, K, dW.data_ptr(), db.data_ptr(), stream), "blvm_linear_dmol_reduce_dw")
                _count()
            grads += [dx, dW, db if bias is not None else None]
            kl_ptrs, klfn_ptrs, kl_chunks = [], [], []
            if L:
                multi = (_lib.KLLevelStruct * L)()
                i = 0
                for li, (lv, (Tz, Z, c)) in enumerate(zip(spec.levels, shapes)):
                    ts = kl_tensors[i:i + lv.n_tensors]
                    i += lv.n_tensors
                    pk = base + 8 * off
                    pf = pk + 8 * B * c
                    off += 2 * B * c
                    d = multi[li]
                    d.lens, d.Tz, d.Z, d.free_nats, d.part_kl, d.part_klfn = _ptr(lv.lens), Tz, Z, lv.free_nats, pk, pf
                    if lv.kind == "kld":
                        gk = torch.empty_like(ts[0]) if need else None
                        d.kl, d.g_kl = ts[0].data_ptr(), _ptr(gk)
                        grads.append(gk)
                    else:
                        n_g = 5 if lv.kind == "mc" else 4
                        g = [torch.empty_like(ts[0]) for _ in range(n_g)] if need else [None] * n_g
                        d.mu_q, d.sd_q, d.mu_p, d.sd_p = ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), ts[3].data_ptr()
                        d.g_mu_q, d.g_sd_q, d.g_mu_p, d.g_sd_p = _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), _ptr(g[3])
                        if lv.kind == "mc":
                            d.z, d.g_z = ts[4].data_ptr(), _ptr(g[4])
                        grads += g
                    kl_ptrs.append(pk)
                    klfn_ptrs.append(pf)
                    kl_chunks.append(c)
                check(lib.blvm_kl_elbo_levels_fwd_grad_scaled(multi, L, B, spec.beta / spec.denom,
                                                              spec.loss_scale.data_ptr() if spec.loss_scale is not None else None, 0, stream),
                      "blvm_kl_elbo_levels_fwd_grad_scaled")
                _count()
            PtrArr, I64Arr = ctypes.c_void_p * max(L, 1), ctypes.c_int64 * max(L, 1)
            check(lib.blvm_elbo_finalize(logp_ptr, chunks, PtrArr(*kl_ptrs), PtrArr(*klfn_ptrs), I64Arr(*kl_chunks), L, x_sl_dev.data_ptr(), B,
                                         spec.beta, spec.denom, rows.data_ptr(), scalars.data_ptr(), _sync_counter(dev).data_ptr(), stream),
                  "blvm_elbo_finalize")
            _count()
        loss = scalars[:1].view(())
        ctx.set_materialize_grads(False)
        ctx.grads = grads
        ctx.prescale = spec.loss_scale
        ctx.consumed = False
        ctx.mark_non_differentiable(scalars, rows, twise)
        _maybe_strict(dev)
        return loss, scalars, rows, twise

    @staticmethod
    def backward(ctx, g_loss, *unused):
        if ctx.consumed:
            raise RuntimeError("blvm_b200 fused ELBO: backward called twice (gradients are produced in the forward pass)")
        ctx.consumed = True
        if g_loss is None:
            return (None, None, None) + tuple(None for _ in ctx.grads)
        g = g_loss if (g_loss.dtype == torch.float64 and g_loss.is_contiguous()) else g_loss.to(torch.float64).contiguous()
        grads = list(ctx.grads)
        unit = _unit_grads.get(g.device.index)
        is_unit = unit is not None and g.data_ptr() == unit.data_ptr()

        def rescale(bufs, factor):
            n = len(bufs)
            with _on_device(g.device):
                check(lib.blvm_scale_inplace_multi((ctypes.c_void_p * n)(*[b.data_ptr() for b in bufs]), (ctypes.c_int64 * n)(*[b.numel() for b in bufs]),
                                                   (ctypes.c_int * n)(*[_DTYPE_CODE[b.dtype] for b in bufs]), n, factor.data_ptr(), _stream()),
                      "blvm_scale_inplace_multi")
            _count()

        head = [b for b in grads[:3] if b is not None]
        rest = [b for b in grads[3:] if b is not None]
        if ctx.prescale is not None:     # all gradients already carry the loss scale S: multiply by grad_output / S (== 1 for scaler.scale(loss))
            if head or rest:
                rescale(head + rest, g / ctx.prescale)
        elif (head or rest) and not is_unit:
            rescale(head + rest, g)
        ctx.grads = None
        return (None, None, None) + tuple(grads)


def fused_linear_elbo_apply(spec: ELBOSpec, y, x_sl_dev, x, weight, bias, kl_tensors: Sequence[torch.Tensor]):
    loss, scalars, rows, twise = _FusedLinearELBO.apply(spec, y, x_sl_dev, x, weight, bias, *kl_tensors)
    return loss.as_subclass(FusedLoss), scalars, rows, twise


def fused_elbo_apply(spec: ELBOSpec, y, x_sl_dev, raw, kl_tensors: Sequence[torch.Tensor]):
    loss, scalars, rows, twise = _FusedELBO.apply(spec, y, x_sl_dev, raw, *kl_tensors)
    return loss.as_subclass(FusedLoss), scalars, rows, twise


# ----------------------------------------------------------------------------------------------------------------------
# fused sample() + mode()
# ----------------------------------------------------------------------------------------------------------------------
def _next_rng(device: torch.device):
    """Philox key / offset taken from (and advanced on) torch's CUDA generator of the device, the way torch's own
    kernels consume it: `torch.manual_seed(s)` makes the sample stream reproducible."""
    if torch.cuda.is_current_stream_capturing():
        # seed and offset are launch ARGUMENTS here: a captured launch would replay the same samples forever and bypass
        # torch's graph-safe generator protocol.  Keep sample()/mode() outside captured regions (the ELBO step itself is
        # capturable: it draws nothing).
        raise RuntimeError("blvm_b200: sample()/mode() cannot be captured in a CUDA graph (the Philox offset is a host-side "
                           "launch argument); call them outside the captured region")
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed = int(gen.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    offset = int(gen.get_offset())
    gen.set_offset(offset + 4)
    return seed, offset // 4 + 1


def dmol_sample_mode(raw: torch.Tensor, K: int, D: int, log_epsilon: float, want_sample=True, want_mode=True):
    """One launch: (sample (*, D), mode (*, D), mode_index (*)) from packed parameters raw (*, K(2D+1)); detached."""
    _require_cuda(raw)
    raw = raw.detach()
    if raw.dtype not in _DTYPE_CODE:
        raw = raw.float()
    raw = raw if raw.is_contiguous() else raw.contiguous()
    batch = raw.shape[:-1]
    N = raw.numel() // raw.shape[-1] if raw.numel() else 0
    sample = torch.empty(*batch, D, dtype=torch.float32, device=raw.device) if want_sample else None
    mode = torch.empty(*batch, D, dtype=torch.float32, device=raw.device) if want_mode else None
    index = torch.empty(batch, dtype=torch.int32, device=raw.device) if want_mode else None
    seed, offset = _next_rng(raw.device)
    with _on_device(raw.device):
        rc = lib.blvm_dmol_sample_mode(_ptr(raw), _DTYPE_CODE[raw.dtype], N, K, D, float(log_epsilon), seed, offset,
                                       _ptr(sample), _ptr(mode), _ptr(index), _stream())
        check(rc, "blvm_dmol_sample_mode")
    _count()
    return sample, mode, index


class _ModeWithGrad(torch.autograd.Function):
    """mode() is differentiable w.r.t. the chosen component's loc in the reference (a gather); the backward scatters."""

    @staticmethod
    def forward(ctx, raw, mode, index, K, D):
        ctx.save_for_backward(index)
        ctx.cfg = (K, D, raw.shape, raw.dtype)
        return mode.to(raw.dtype) if raw.dtype != torch.float32 else mode

    @staticmethod
    def backward(ctx, g):
        (index,) = ctx.saved_tensors
        K, D, shape, dtype = ctx.cfg
        graw = torch.zeros(shape, dtype=dtype, device=g.device)
        cols = (K + index.long().unsqueeze(-1) + 2 * K * torch.arange(D, device=g.device)).reshape(*shape[:-1], D)
        graw.scatter_(-1, cols, g.to(dtype))
        return graw, None, None, None, None


def mode_with_grad(raw, mode, index, K, D):
    return _ModeWithGrad.apply(raw, mode, index, K, D)


# ----------------------------------------------------------------------------------------------------------------------
# Quantize
# ----------------------------------------------------------------------------------------------------------------------
def quantize_indices(x: torch.Tensor, boundaries: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, boundaries)
    x = _as_f32c(x)
    boundaries = _as_f32c(boundaries)
    out = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    with _on_device(x.device):
        check(lib.blvm_quantize(_ptr(x), x.numel(), _ptr(boundaries), boundaries.numel(), _ptr(out), _stream()), "blvm_quantize")
    _count()
    return out

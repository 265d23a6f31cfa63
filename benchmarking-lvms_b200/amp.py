"""Loss-scale awareness for fp16 parameters (the reference's default `--use_amp True`: fp16 autocast + GradScaler,
experiments/experiment_*_audio.py: `scaler.scale(loss).backward()`).

Gradients of the path are ~1/sum(x_sl) ~ 1e-7, below fp16's range, so they can only be written in fp16 AFTER the
GradScaler's factor has been applied.  Without knowing the scaler the fused op therefore defers the gradient launch to
backward (value kernel in forward, value + gradient kernel in backward: 93 + 127 us at B=256, T=16000, K=10).  When the
scaler is known, the forward kernel reads its scale from the device (`scaler._scale`, no host sync) and writes
scale x gradient in the same pass (127 us); backward then only has to multiply by `grad_output / scale`, which is exactly 1
for `scaler.scale(loss).backward()` (one early-exit launch).

A scaler becomes known either explicitly (`fused_elbo(..., grad_scaler=scaler)` / `register_grad_scaler(scaler)`) or, after
`patch_blvm()`, automatically: GradScaler construction is observed and the single enabled instance is used.
"""
import warnings
import weakref

import torch

__all__ = ["register_grad_scaler", "active_grad_scaler", "observe_grad_scalers"]

_scalers = weakref.WeakSet()
_observed = []


def register_grad_scaler(scaler):
    """Make `scaler` (torch.amp.GradScaler / torch.cuda.amp.GradScaler) known to the fused op."""
    _scalers.add(scaler)
    return scaler


_warned_ambiguous = False


def active_grad_scaler(device: torch.device):
    """The one enabled, registered scaler whose scale tensor already lives on `device`, else None (ambiguity, a disabled
    scaler or a scale that has not been created yet all fall back to the deferred-gradient path)."""
    found = None
    for s in list(_scalers):
        if not getattr(s, "_enabled", False):
            continue
        scale = getattr(s, "_scale", None)
        if scale is None and hasattr(s, "_lazy_init_scale_growth_tracker"):
            # the scale tensor is created by the first `scaler.scale(loss)`, i.e. AFTER the first forward pass; create it now
            # exactly as that call would (same init value, same device) so that the very first step is single-pass too
            s._lazy_init_scale_growth_tracker(device)
            scale = s._scale
        if scale is None or scale.device != device:
            continue
        if found is not None:
            global _warned_ambiguous
            if not _warned_ambiguous:
                _warned_ambiguous = True
                warnings.warn("blvm_b200: more than one enabled GradScaler is alive on this device, so the loss scale of the step is "
                              "ambiguous: fp16 gradients fall back to the deferred (two-pass) path and the fused likelihood head is not "
                              "taken.  Pass fused_elbo(..., grad_scaler=scaler) or drop the stale scaler objects.", RuntimeWarning, stacklevel=3)
            return None
        found = s
    return found


def ensure_scale(scaler, device: torch.device) -> bool:
    """True if `scaler` is enabled and has its device-side scale on `device`; creates the scale tensor exactly as the first
    `scaler.scale(loss)` call would if that has not happened yet (so that the very first step is single-pass too)."""
    if scaler is None or not getattr(scaler, "_enabled", False):
        return False
    if getattr(scaler, "_scale", None) is None and hasattr(scaler, "_lazy_init_scale_growth_tracker"):
        scaler._lazy_init_scale_growth_tracker(device)
    scale = getattr(scaler, "_scale", None)
    return scale is not None and scale.device == device


def scale_tensor_f64(scaler) -> torch.Tensor:
    """A private fp64 copy of the scaler's current scale (one tiny kernel; `update()` later mutates the original)."""
    return scaler._scale.detach().to(torch.float64).reshape(())


def observe_grad_scalers():
    """Register every GradScaler constructed from now on (used by patch_blvm so that experiment scripts run unchanged)."""
    if _observed:
        return
    classes = []
    for mod, name in (("torch.amp.grad_scaler", "GradScaler"), ("torch.cuda.amp.grad_scaler", "GradScaler")):
        try:
            cls = getattr(__import__(mod, fromlist=[name]), name)
        except Exception:
            continue
        if cls not in classes:
            classes.append(cls)
    for cls in classes:
        orig = cls.__init__
        if getattr(orig, "_blvm_observed", False):
            continue

        def init(self, *a, __orig=orig, **k):
            __orig(self, *a, **k)
            _scalers.add(self)

        init._blvm_observed = True
        cls.__init__ = init
        _observed.append((cls, orig))


def stop_observing():
    while _observed:
        cls, orig = _observed.pop()
        cls.__init__ = orig

"""`patch_blvm()`: swap the kernels into an importable reference tree so that `experiments/experiment_*_audio.py` run
unchanged.  The reference binds names with `from ... import ...` (vrnn.py:26-27, srnn.py:24-25,
stcn.py:28-29, clockwork_vae.py:27-28), so every `blvm.*` module's globals are rebound, not only the defining module."""
import sys
from types import ModuleType

from . import amp, distributions, elbo, log_likelihoods, metrics, variational

__all__ = ["patch_blvm", "unpatch_blvm"]

_saved = []


def _rebind_everywhere(original, replacement):
    for name, module in list(sys.modules.items()):
        if not isinstance(module, ModuleType) or not (name == "blvm" or name.startswith("blvm.")):
            continue
        for attr, value in list(vars(module).items()):
            if value is original:
                _saved.append((module, attr, original))
                setattr(module, attr, replacement)


def _rebind_method(cls, name, replacement):
    _saved.append((cls, name, cls.__dict__[name]))
    setattr(cls, name, replacement)


def patch_blvm(fuse_linear: bool = False, lazy_samples: bool = False):
    """Import `blvm` (must be on sys.path) and rebind the hot-path functions, classes and model reducers to blvm_b200.
    Returns the list of `module.attr` names that were rebound.

    fuse_linear=True: every DiscretizedLogisticMixtureDense constructed from now on also fuses its `nn.Linear` into the
    likelihood kernel when it runs under AMP (fp16 / bf16 activations, num_mix = 10, even x_dim <= 79): Linear, DMoL value +
    gradient and the Linear's backward as one tcgen05 tensor-core kernel (csrc/linear_dmol_kernel.cuh).

    lazy_samples=True: `likelihood.sample(params)` / `.mode(params)` of those modules return promises that launch the fused sample + mode
    kernel only when somebody reads them (the models call both on every training step, vrnn.py:332-333, and a training loop never looks):
    a patched training step is then 3 launches of ours -- likelihood, KL, finalize -- and draws nothing (so it can be captured in a CUDA
    graph).  The random draw happens at the first read; a lazily read mode is detached."""
    import importlib

    import blvm.modules.distributions as ref_dist
    import blvm.utils.log_likelihoods as ref_ll
    import blvm.utils.variational as ref_var
    import blvm.models  # noqa: F401  (populates sys.modules with the model modules)

    before = len(_saved)
    _rebind_everywhere(ref_ll.discretized_logistic_mixture_ll, log_likelihoods.discretized_logistic_mixture_ll)
    _rebind_everywhere(ref_ll.discretized_logistic_ll, log_likelihoods.discretized_logistic_ll)
    _rebind_everywhere(ref_ll.gaussian_mixture_ll, log_likelihoods.gaussian_mixture_ll)
    _rebind_everywhere(ref_var.kl_divergence_gaussian, variational.kl_divergence_gaussian)
    _rebind_everywhere(ref_dist.DiagonalGaussianMixtureDense, distributions.DiagonalGaussianMixtureDense)
    # gaussian_ll / DiagonalGaussianDense are NOT rebound: the model bodies use them for the latent layers
    # (vrnn.py:81,91), which are outside this path; blvm_b200.gaussian_ll / kl_divergence_gaussian_mc exist for callers
    # that want the kernels (e.g. bottom-up STCN).
    dmol_cls = distributions.DiscretizedLogisticMixtureDense
    if fuse_linear or lazy_samples:
        class DiscretizedLogisticMixtureDense(distributions.DiscretizedLogisticMixtureDense):   # same name: repr / checkpoints unchanged
            def __init__(self, *args, **kwargs):
                if fuse_linear:
                    kwargs.setdefault("fuse_linear", True)
                if lazy_samples:
                    kwargs.setdefault("lazy_samples", True)
                super().__init__(*args, **kwargs)
        dmol_cls = DiscretizedLogisticMixtureDense
    _rebind_everywhere(ref_dist.DiscretizedLogisticMixtureDense, dmol_cls)
    _rebind_everywhere(ref_dist.DiscretizedLogisticDense, distributions.DiscretizedLogisticDense)

    vrnn = importlib.import_module("blvm.models.vrnn")
    srnn = importlib.import_module("blvm.models.srnn")
    stcn = importlib.import_module("blvm.models.stcn.stcn")
    cwvae = importlib.import_module("blvm.models.clockwork_vae.clockwork_vae")
    wavenet = importlib.import_module("blvm.models.wavenet.wavenet")
    _rebind_method(vrnn.VRNN, "compute_elbo", elbo.vrnn_compute_elbo)
    _rebind_method(srnn.SRNN, "compute_elbo", elbo.srnn_compute_elbo)
    _rebind_method(cwvae.CWVAE, "compute_elbo", elbo.cwvae_compute_elbo)
    _rebind_method(wavenet.WaveNet, "compute_loss", elbo.wavenet_compute_loss)

    _rebind_method(stcn.STCN, "compute_loss", elbo.stcn_compute_loss)   # top-down (analytic KL) and bottom-up (MC KL)
    # `--use_amp True` (fp16 autocast + GradScaler, the reference's benchmark default): learn about the script's scaler so
    # that fp16 gradients are produced in one pass with its device-side scale (amp.py)
    amp.observe_grad_scalers()
    # the models' Metric objects (vrnn.py:346-355 ...) read their device values lazily: one D2H sync per step instead of
    # one per metric (metrics.py:241-244)
    import blvm.evaluation.metrics as ref_metrics
    metrics.patch_metric_syncs(ref_metrics)
    return [f"{getattr(o, '__name__', o)}.{a}" for o, a, _ in _saved[before:]]


def unpatch_blvm():
    amp.stop_observing()
    metrics.unpatch_metric_syncs()
    while _saved:
        owner, attr, original = _saved.pop()
        setattr(owner, attr, original)

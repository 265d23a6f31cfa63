"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/blvm_b200.h declares
(no compute calls without a GPU), the Python mirror keeps the reference's API surface, and patch_blvm() rebinds an
importable reference tree (skipped where the reference is absent, e.g. on the GPU box)."""
import ctypes
import inspect
import os
import re
import sys

import pytest
import torch

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "blvm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(blvm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import blvm_b200
    lib = ctypes.CDLL(blvm_b200.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/blvm_b200.h but not exported by {blvm_b200.LIB_PATH}"
    assert set(names) == set(blvm_b200._lib.SIGNATURES), "ctypes signatures and header out of sync"
    assert blvm_b200._lib.lib.blvm_version() == 201
    # host-only entry points are callable without a GPU
    assert blvm_b200._lib.lib.blvm_dmol_chunks(16000, 10, 1) == 125
    assert blvm_b200._lib.lib.blvm_dmol_chunks(16000, 1, 1) == 16      # 8 samples per thread at K = 1
    assert blvm_b200._lib.lib.blvm_dmol_chunks(16000, 7, 1) == 125     # generic kernel tile
    assert blvm_b200._lib.lib.blvm_kl_chunks(250 * 64) == 16


def test_library_has_sm100a_code_and_tma():
    """The shipped .so contains sm_100a SASS with the TMA bulk-copy instruction (UBLKCP)."""
    import shutil
    import subprocess

    import blvm_b200
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", blvm_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    syms = subprocess.run(["cuobjdump", "-elf", blvm_b200.LIB_PATH], capture_output=True, text=True).stdout
    import re as _re
    names = sorted(set(_re.findall(r"_ZN4blvm16dmol_tile_kernelILi10ELi128ELb1ELi0EfLi0EE\w*DmolArgsE\b", syms)))
    assert names, "K=10 fp32 fwd+grad tile kernel not found in the library"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], blvm_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_module_api_surface_matches_reference():
    import blvm_b200 as B
    m = B.DiscretizedLogisticMixtureDense(x_dim=30, y_dim=1, num_mix=10, num_bins=2 ** 16)
    assert (m.x_dim, m.y_dim, m.num_mix, m.num_bins, m.log_epsilon, m.out_features) == (30, 1, 10, 65536, -7.0, 30)
    assert list(m.state_dict().keys()) == ["params.weight", "params.bias"]           # distributions.py:345
    with pytest.raises(NotImplementedError):
        m.get_distribution(None)                                                      # distributions.py:352-354
    p = m(torch.randn(2, 5, 30))
    assert len(p) == 3 and p[0].shape == (2, 5, 10) and p[1].shape == (2, 5, 1, 10) and p[2].shape == (2, 5, 1, 10)
    assert float(p[2].min()) >= -7.0                                                 # clamp, distributions.py:386
    lp, lc, ls = p                                                                    # iterable like the tuple
    assert m.rsample(p).shape == (2, 5, 1)                                            # torch-op composition, any device
    for call in (m.mode, m.sample):                                                   # kernels: CPU tensors raise, no torch fallback
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call(p)
    assert float(m.rsample(p).abs().max()) <= 1.0
    d = B.DiscretizedLogisticDense(x_dim=8, y_dim=1, num_bins=256)
    assert d.out_features == 2 and list(d.state_dict().keys()) == ["params.weight", "params.bias"]
    q = d(torch.randn(3, 8))
    assert q[0].shape == (3, 1) and q[1].shape == (3, 1)
    # functions keep the reference's parameter names and defaults
    sig = inspect.signature(B.discretized_logistic_mixture_ll)
    assert list(sig.parameters) == ["y", "logit_probs", "locs", "log_scales", "num_bins", "reduce_dim"]
    assert sig.parameters["num_bins"].default == 256 and sig.parameters["reduce_dim"].default == -1
    assert list(inspect.signature(B.kl_divergence_gaussian).parameters) == ["mu_q", "sd_q", "mu_p", "sd_p"]
    assert list(inspect.signature(B.discount_free_nats).parameters) == ["kld", "free_nats", "shared_dims"]
    assert list(inspect.signature(B.vrnn_compute_elbo).parameters) == ["self", "y", "parameters", "kld_twise", "x_sl", "stride", "beta", "free_nats"]
    assert list(inspect.signature(B.cwvae_compute_elbo).parameters) == ["self", "y", "seq_mask", "level_masks", "x_sl", "parameters", "kld_layerwise", "beta", "free_nats"]
    assert list(inspect.signature(B.stcn_compute_loss).parameters) == ["self", "y", "x_sl", "parameters", "mu_p", "sd_p", "mu_q", "sd_q", "z", "free_nats", "beta"]
    assert list(inspect.signature(B.wavenet_compute_loss).parameters) == ["self", "y", "x_sl", "parameters"]


def test_host_side_helpers():
    import blvm_b200 as B
    x_sl = torch.tensor([96, 50, 33, 8])
    m = B.sequence_mask(x_sl, dtype=float)
    assert m.dtype == torch.float64 and m.shape == (4, 96) and m.sum(1).tolist() == [96, 50, 33, 8]
    # ceil(x_sl / stride) == mask[:, ::stride].sum(1)  (vrnn.py:271)
    for s in (1, 3, 8, 64):
        assert torch.equal(B.level_lengths(x_sl, s), B.sequence_mask(x_sl)[:, ::s].sum(1))
    kl = torch.rand(2, 3, 4)
    assert B.discount_free_nats(kl, 0) is kl and B.discount_free_nats(kl, None) is kl   # variational.py:107-108
    fn = B.discount_free_nats(kl, 2.0, shared_dims=-1)
    assert torch.equal(fn, torch.maximum(kl, torch.tensor(0.5)))
    rm = B.RunningMean()
    rm.update(1.0, 10)
    assert rm.update(3.0, 30) == pytest.approx(2.5)                                   # metrics.py:253-264


def test_no_cpu_fallback_and_oracle_not_imported_by_product():
    import blvm_b200 as B
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.kl_divergence_gaussian(torch.zeros(2), torch.ones(2), torch.zeros(2), torch.ones(2))
    pkg = os.path.join(ROOT, "benchmarking-lvms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "blvm_oracle" not in src, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/blvm"), reason="reference tree not present (GPU box)")
def test_patch_blvm_rebinds_reference_names():
    os.environ.setdefault("BLVM_DATA_ROOT_DIRECTORY", "/tmp/blvmdata")
    os.makedirs(os.environ["BLVM_DATA_ROOT_DIRECTORY"], exist_ok=True)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "_ref_shims"))
    sys.path.insert(0, "/root/reference")
    import blvm_b200 as B
    try:
        names = B.patch_blvm()
        import importlib
        vrnn = importlib.import_module("blvm.models.vrnn")
        stcn = importlib.import_module("blvm.models.stcn.stcn")
        dist_mod = importlib.import_module("blvm.modules.distributions")
        assert vrnn.kl_divergence_gaussian is B.kl_divergence_gaussian          # vrnn.py:26-27 bound by name
        assert vrnn.DiscretizedLogisticMixtureDense is B.DiscretizedLogisticMixtureDense
        assert dist_mod.discretized_logistic_mixture_ll is B.discretized_logistic_mixture_ll
        assert vrnn.VRNN.compute_elbo is B.vrnn_compute_elbo
        assert stcn.STCN.compute_loss is not None and len(names) >= 10
        # a reference model constructed after patching owns our likelihood module (checkpoint keys unchanged)
        model = vrnn.VRNNAudio(input_size=200, hidden_size=32, latent_size=8, likelihood="DMoL")
        assert isinstance(model.vrnn.likelihood, B.DiscretizedLogisticMixtureDense)
        assert any(k.endswith("likelihood.params.weight") for k in model.state_dict())
    finally:
        B.unpatch_blvm()
    vrnn = sys.modules["blvm.models.vrnn"]
    assert vrnn.VRNN.compute_elbo is not B.vrnn_compute_elbo


def test_grad_scaler_observation_and_numa_binding_are_safe_without_a_gpu():
    """amp.observe_grad_scalers() registers scalers constructed afterwards and restores GradScaler.__init__ when stopped; a
    disabled scaler (or one whose scale tensor does not exist yet) is never 'active'; the NUMA helper is a no-op when the
    topology cannot be read."""
    import torch
    import blvm_b200
    from blvm_b200 import amp
    orig_init = torch.amp.GradScaler.__init__
    amp.observe_grad_scalers()
    try:
        assert torch.amp.GradScaler.__init__ is not orig_init
        s = torch.amp.GradScaler("cuda", enabled=False)
        assert s in amp._scalers
        assert amp.active_grad_scaler(torch.device("cuda", 0)) is None
    finally:
        amp.stop_observing()
    assert torch.amp.GradScaler.__init__ is orig_init
    s2 = torch.amp.GradScaler("cuda", enabled=False)
    assert s2 not in amp._scalers
    assert blvm_b200.register_grad_scaler(s2) is s2 and s2 in amp._scalers
    amp._scalers.discard(s2)
    amp._scalers.discard(s)
    if not torch.cuda.is_available():
        assert blvm_b200.bind_to_gpu_numa_node(0) is None

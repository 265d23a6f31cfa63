"""The drop-in claim, on the GPU, with the REAL reference models: every audio model of the reference (VRNN, SRNN,
Clockwork-VAE, STCN top-down and bottom-up, WaveNet, LSTM) is built from the staged, unmodified reference
(`oracle/_ref`, written by oracle/make_ref.py) on cuda:0 and taken through the training step of
`experiments/experiment_vrnn_audio.py:216-232`

    with autocast(enabled=use_amp): loss, metrics, outputs = model(x, x_sl, beta=..., free_nats=...)
    optimizer.zero_grad(set_to_none=True); scaler.scale(loss).backward(); scaler.unscale_(optimizer)
    clip_grad_value_; clip_grad_norm_; scaler.step(optimizer); scaler.update(); read the metrics

once unpatched (the reference's own eager op chain) and once under `blvm_b200.patch_blvm()` (the kernels), from the same
weights, inputs and RNG state.

Tolerances and anchor.  A whole model cannot be rerun in fp64 with the same latent samples (the RNG stream depends on the
dtype), so the anchor is the unpatched reference model with ONLY the two functions of the path evaluated in fp64
(`discretized_logistic_mixture_ll` and `kl_divergence_gaussian` called with `.double()` inputs at the reference's own call
sites; model body, masks and reductions untouched, same weights / inputs / RNG).  That removes the two cancellation-prone
expressions of the reference's fp32 run -- `sigmoid(a) - sigmoid(b)` on 16-bit audio (tests/parity.py) and
`log sd_p - log sd_q + ... - 1/2` for q ~ p (an untrained Clockwork-VAE: the reference's fp32 run is 3e-3 of a tensor's scale
away from the anchor on the top level's weights).  Against that anchor: loss rel 2e-6, every weight gradient within 1e-4 of its tensor's
scale, and never further away than the reference's own fp32 run is (the kernel-level tests hold the 1e-5 / 1e-6 bars).
"""
import copy
import importlib
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

MODELS = ["vrnn", "srnn", "cwvae", "stcn", "stcn_bottom_up", "wavenet", "lstm"]
LOSS_RTOL = 2e-6
GRAD_TOL = 1e-4        # of the tensor's max |gradient|


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        pytest.skip(ref_loader.why_unavailable())
    ref_loader.load()
    import blvm.models as M
    return M


def build(M, name):
    torch.manual_seed(3)
    lik = importlib.import_module("blvm.modules.distributions").DiscretizedLogisticMixtureDense
    if name == "vrnn":
        return M.VRNNAudio(input_size=200, hidden_size=64, latent_size=16, likelihood="DMoL")
    if name == "srnn":
        return M.SRNNAudio(likelihood="DMoL", input_size=64, hidden_size=64, latent_size=16, num_bins=2 ** 16)
    if name == "cwvae":
        return M.CWVAEAudio(z_size=[16, 8], h_size=[32, 32], strides=[16, 4], num_level_layers=2, stride_per_layer=4,
                            likelihood="DMoL", num_bins=2 ** 16)
    if name == "stcn":
        return M.STCN(likelihood="DMoL", n_layers=2, latent_size=[16, 8], res_channels=32)
    if name == "stcn_bottom_up":   # Monte-Carlo KL at the sampled z (stcn.py:288)
        return M.STCN(likelihood="DMoL", n_layers=2, latent_size=[16, 8], res_channels=32, top_down=False)
    if name == "wavenet":
        return M.WaveNet(likelihood=lik(x_dim=32, y_dim=1, num_mix=10, num_bins=2 ** 16), n_layers=3, n_stacks=1, res_channels=32)
    if name == "lstm":
        return M.LSTMAudio(stack_size=64, hidden_size=32, num_bins=2 ** 16)
    raise KeyError(name)


def inputs(name):
    T = 3200 if name != "cwvae" else 2048
    g = torch.Generator().manual_seed(1)
    x = (torch.randint(0, 65536, (4, T), generator=g).float() / 65535 * 2 - 1)
    x_sl = torch.tensor([T, T - 37 * 8, T - 640, T // 2])   # padded utterances
    kwargs = {} if name in ("wavenet", "lstm") else dict(beta=0.5, free_nats=0.25)
    return x, x_sl, kwargs


def train_step(model, x, x_sl, kwargs, use_amp, scaler_cls):
    """experiments/experiment_vrnn_audio.py:198,216-232, restated."""
    optimizer = torch.optim.SGD(model.parameters(), lr=0.0)   # lr 0: the step runs, the weights stay comparable
    scaler = scaler_cls(enabled=use_amp)
    model.train()
    torch.manual_seed(5)   # latent samples / sample() inside forward consume RNG: same stream for both runs
    with torch.autocast("cuda", dtype=torch.float16, enabled=use_amp):
        loss, metrics, outputs = model(x.cuda(), x_sl, **kwargs)
    optimizer.zero_grad(set_to_none=True)
    scaler.scale(loss).backward()
    scaler.unscale_(optimizer)
    grads = {n: p.grad.detach().double().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}
    torch.nn.utils.clip_grad_value_(model.parameters(), 3000.0)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 3000.0)
    scaler.step(optimizer)
    scaler.update()
    values = [(m.name, float(m.value)) for m in metrics]
    return loss.detach(), grads, values, outputs, scaler


class likelihood_in_fp64:
    """Context: the reference's DMoL log-likelihood and Gaussian-KL FUNCTIONS evaluate in fp64 (inputs upcast, result cast
    back to fp32) wherever the reference models call them; everything else is the unmodified fp32 model."""

    def __init__(self, ulp_noise: float = 0.0):
        # ulp_noise > 0: every per-sample log-prob is multiplied by (1 + ulp_noise * n_t), n_t ~ N(0, 1) fixed: a relative
        # perturbation of the per-sample likelihood gradients by about one fp32 ulp -- the sensitivity probe of the test below
        self.ulp_noise = ulp_noise

    def __enter__(self):
        self.saved = []
        dist = importlib.import_module("blvm.modules.distributions")
        var = importlib.import_module("blvm.utils.variational")
        ll, kl = dist.discretized_logistic_mixture_ll, var.kl_divergence_gaussian
        ulp_noise = self.ulp_noise

        def ll64(y, logit_probs, locs, log_scales, **kw):
            lp = ll(y.double(), logit_probs.double(), locs.double(), log_scales.double(), **kw)
            if ulp_noise:
                g = torch.Generator(device=lp.device).manual_seed(77)
                lp = lp * (1.0 + ulp_noise * torch.randn(lp.shape, generator=g, device=lp.device, dtype=lp.dtype))
            return lp.to(logit_probs.dtype)

        def kl64(mu_q, sd_q, mu_p, sd_p):
            return kl(mu_q.double(), sd_q.double(), mu_p.double(), sd_p.double()).to(mu_q.dtype)

        for name, module in list(sys.modules.items()):
            if name == "blvm" or name.startswith("blvm."):
                for attr, value in list(vars(module).items()):
                    if value is ll or value is kl:
                        self.saved.append((module, attr, value))
                        setattr(module, attr, ll64 if value is ll else kl64)
        return self

    def __exit__(self, *exc):
        for module, attr, value in self.saved:
            setattr(module, attr, value)
        return False


def run_pair(M, name, use_amp, anchor64=False, fuse_linear=False):
    import blvm_b200 as B
    x, x_sl, kwargs = inputs(name)
    ref_model = build(M, name).cuda()
    state = copy.deepcopy(ref_model.state_dict())
    scaler_cls = torch.cuda.amp.GradScaler
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = train_step(ref_model, x, x_sl, kwargs, use_amp, scaler_cls)
        if anchor64:
            with likelihood_in_fp64():
                r64 = train_step(ref_model, x, x_sl, kwargs, use_amp, scaler_cls)
                r64b = train_step(ref_model, x, x_sl, kwargs, use_amp, scaler_cls)
            with likelihood_in_fp64(ulp_noise=1.2e-7):
                r64n = train_step(ref_model, x, x_sl, kwargs, use_amp, scaler_cls)
            r = (r, r64, r64b, r64n)
    del ref_model
    try:
        rebound = B.patch_blvm(fuse_linear=fuse_linear)
        assert rebound, "patch_blvm() rebound nothing"
        model = build(M, name).cuda()
        assert set(model.state_dict()) == set(state)                      # checkpoint keys unchanged
        model.load_state_dict(state)
        assert any(isinstance(m, B.DiscretizedLogisticMixtureDense) for m in model.modules())
        B.reset_launch_count()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = train_step(model, x, x_sl, kwargs, use_amp, torch.cuda.amp.GradScaler)   # constructed under the patch: observed
        launches = B.launch_count()
        B.check_input_range()
    finally:
        B.unpatch_blvm()
    return r, o, launches


@pytest.mark.parametrize("name", MODELS)
def test_patched_training_step_matches_reference_fp32(name, ref):
    (ref32, ref64, ref64b, ref64n), (loss_o, grads_o, vals_o, out_o, _), launches = run_pair(ref, name, use_amp=False, anchor64=True)
    loss_32, grads_32, grads_64b, grads_64n = ref32[0], ref32[1], ref64b[1], ref64n[1]
    loss_r, grads_r, vals_r, out_r, _ = ref64          # the anchor: reference model, likelihood function in fp64
    assert loss_o.dtype == loss_r.dtype, (loss_o.dtype, loss_r.dtype)     # float64 for VRNN/SRNN, float32 elsewhere
    rel = abs(float(loss_o) - float(loss_r)) / abs(float(loss_r))
    rel32 = abs(float(loss_32) - float(loss_r)) / abs(float(loss_r))
    assert set(grads_o) == set(grads_r), "a parameter lost (or gained) its gradient under the patch"
    # Per tensor: our distance from the anchor, the reference fp32 run's distance from it, the anchor's own run-to-run noise
    # (the same anchor step executed twice: cuDNN / index-add backward kernels accumulate with atomics) and its SENSITIVITY:
    # how far the tensor moves when every per-sample log-prob of the anchor is perturbed by one fp32 ulp (1.2e-7 relative).
    # Some models amplify that by 1e3-1e4 (an untrained Clockwork-VAE: weight gradients that are small residuals of large
    # cancelling terms); no fp32 implementation of the likelihood can then agree with the anchor better than the probe does.
    # A tensor passes within GRAD_TOL of its scale, or within 3x the anchor's own noise, or within 8x its one-ulp sensitivity
    # (the kernels' per-sample values and gradients are good to 2-4 fp32 ulps, tests/parity.py).
    worst, worst_name, worst32, rows = 0.0, None, 0.0, []
    for n, g in grads_r.items():
        scale = float(g.abs().max())
        if scale == 0.0:
            assert float(grads_o[n].abs().max()) == 0.0, n
            continue
        e = float((grads_o[n] - g).abs().max()) / scale
        e32 = float((grads_32[n] - g).abs().max()) / scale
        noise = float((grads_64b[n] - g).abs().max()) / scale
        sens = float((grads_64n[n] - g).abs().max()) / scale
        rows.append((e, e32, noise, sens, n))
        worst32 = max(worst32, e32)
        if e > worst:
            worst, worst_name = e, n
    rows.sort(reverse=True)
    print(f"\n[{name}] loss anchor {float(loss_r):.8f} ours {float(loss_o):.8f} rel {rel:.2e} (reference fp32: {rel32:.2e}); worst "
          f"weight-grad error {worst:.2e} of its scale ({worst_name}) (reference fp32: {worst32:.2e}); {len(grads_r)} gradient "
          f"tensors; {launches} blvm launches")
    for e, e32, noise, sens, n in rows[:4] + ([r for r in rows[4:] if r[0] > GRAD_TOL / 10] if os.environ.get("BLVM_TEST_VERBOSE") else []):
        print(f"[{name}]    {n}: ours-anchor {e:.2e}, reference fp32-anchor {e32:.2e}, anchor run-to-run {noise:.2e}, one-ulp sensitivity {sens:.2e}")
    assert rel < LOSS_RTOL
    bad = [(e, noise, sens, n) for e, e32, noise, sens, n in rows if e > max(GRAD_TOL, 3 * noise, 8 * sens)]
    assert not bad, bad[:3]
    assert [n for n, _ in vals_o] == [n for n, _ in vals_r]              # metric list order preserved (vrnn.py:346-355)
    for (n, a), (_, b) in zip(vals_o, vals_r):
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7, err_msg=f"metric {n}")
    for key in ("elbo", "log_prob", "kl", "kld"):
        if hasattr(out_r, key):
            a, b = getattr(out_o, key), getattr(out_r, key)
            assert a.dtype == b.dtype and a.shape == b.shape, key
            np.testing.assert_allclose(a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy(), rtol=2e-5, err_msg=key)
    assert sorted(vars(out_o)) == sorted(vars(out_r))


@pytest.mark.parametrize("name", ["vrnn", "srnn", "cwvae", "stcn", "wavenet"])
def test_patched_amp_training_step(name, ref, monkeypatch):
    """`--use_amp True` (the reference's benchmark default, experiments/benchmarks.txt): fp16 autocast + GradScaler.  The
    scaler the loop constructs is observed by patch_blvm(), so the fp16 likelihood gradient is written in the forward pass,
    pre-multiplied by the device-side scale: no deferred value+gradient launch in backward."""
    from blvm_b200 import amp, ops
    calls = {"n": 0}
    real = ops._dmol_call
    monkeypatch.setattr(ops, "_dmol_call", lambda *a, **k: (calls.__setitem__("n", calls["n"] + 1), real(*a, **k))[1])
    (loss_r, grads_r, vals_r, _, _), (loss_o, grads_o, vals_o, _, scaler), launches = run_pair(ref, name, use_amp=True)
    assert calls["n"] == 0, "the fp16 gradient was deferred to backward: the GradScaler was not observed"
    assert scaler in amp._scalers
    rel = abs(float(loss_o) - float(loss_r)) / abs(float(loss_r))
    worst = 0.0
    for n, g in grads_r.items():
        scale = float(g.abs().max())
        if scale > 0 and torch.isfinite(g).all():
            worst = max(worst, float((grads_o[n] - g).abs().max()) / scale)
    print(f"\n[{name} amp] loss ref {float(loss_r):.6f} ours {float(loss_o):.6f} rel {rel:.2e}; worst weight-grad error {worst:.2e}; "
          f"{launches} blvm launches")
    assert rel < 2e-3          # fp16 body: the two runs differ by fp16 rounding of the Linear output's consumers
    assert worst < 5e-2
    assert set(grads_o) == set(grads_r)


@pytest.mark.parametrize("name", ["vrnn", "srnn", "stcn"])
def test_patched_amp_step_with_the_fused_tensor_core_head(name, ref, monkeypatch):
    """`patch_blvm(fuse_linear=True)` under `--use_amp True`: the likelihood's nn.Linear, the DMoL value + gradient and the Linear's
    backward run as ONE tcgen05 kernel (csrc/linear_dmol_kernel.cuh); the (B, T, 3K) parameter tensor is never built because
    `sample()` / `mode()` of the unevaluated parameters are promises nobody reads during a training step."""
    from blvm_b200 import ops
    from blvm_b200.variational import LazyResult
    calls = {"head": 0}
    real = ops.fused_linear_elbo_apply
    monkeypatch.setattr(ops, "fused_linear_elbo_apply", lambda *a, **k: (calls.__setitem__("head", calls["head"] + 1), real(*a, **k))[1])
    (loss_r, grads_r, vals_r, _, _), (loss_o, grads_o, vals_o, out_o, _), launches = run_pair(ref, name, use_amp=True, fuse_linear=True)
    assert calls["head"] == 1, "the fused head was not taken"
    rec = out_o.reconstructions
    assert isinstance(rec, LazyResult) and rec._blvm_value is None          # promised, never evaluated by the training step
    rel = abs(float(loss_o) - float(loss_r)) / abs(float(loss_r))
    worst = 0.0
    for n, g in grads_r.items():
        scale = float(g.abs().max())
        if scale > 0 and torch.isfinite(g).all():
            worst = max(worst, float((grads_o[n] - g).abs().max()) / scale)
    print(f"\n[{name} amp fused head] loss ref {float(loss_r):.6f} ours {float(loss_o):.6f} rel {rel:.2e}; worst weight-grad error {worst:.2e}; "
          f"{launches} blvm launches")
    assert rel < 2e-3 and worst < 5e-2 and set(grads_o) == set(grads_r)
    assert tuple(rec.shape) == tuple(out_o.y.shape[:2]) + (1,) and float(rec.abs().max()) <= 1.0 and rec._blvm_value is not None   # reading works


@pytest.mark.parametrize("name", ["vrnn", "srnn", "cwvae"])
def test_patched_step_has_one_fused_kl_launch_and_one_metric_sync(name, ref):
    """under the patch the latent levels go through ONE fused KL launch (the elementwise KL
    is never materialised: no blvm_kl_gaussian_fwd / _bwd, no kl_reduce) and the model's Metric objects cost ONE
    device->host synchronisation per step instead of one each."""
    import blvm_b200 as B
    from blvm_b200 import ops
    M = ref
    x, x_sl, kwargs = inputs(name)
    model_ref = build(M, name).cuda()
    state = copy.deepcopy(model_ref.state_dict())

    def forward_and_read(model):
        torch.manual_seed(5)
        torch.cuda.synchronize()
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            torch.cuda.set_sync_debug_mode("warn")
            try:
                loss, metrics, _ = model(x.cuda(), x_sl, **kwargs)
                vals = [m.value for m in metrics]
            finally:
                torch.cuda.set_sync_debug_mode("default")
        found = [(os.path.relpath(i.filename, ROOT) if i.filename.startswith(ROOT) else i.filename, i.lineno) for i in w
                 if "synchroniz" in str(i.message).lower()]
        return found, vals

    syncs_ref, _ = forward_and_read(model_ref)
    try:
        B.patch_blvm()
        model = build(M, name).cuda()
        model.load_state_dict(state)
        eager = {"n": 0}
        real = ops.kl_gaussian
        ops.kl_gaussian = lambda *a: (eager.__setitem__("n", eager["n"] + 1), real(*a))[1]
        try:
            B.reset_launch_count()
            syncs, _ = forward_and_read(model)
            launches = B.launch_count()
        finally:
            ops.kl_gaussian = real
    finally:
        B.unpatch_blvm()
    def ours(found):   # synchronisations caused by the path (likelihood / KL / compute_elbo / Metric objects), not by the model body
        # (sequence_mask, operations.py:90-119, belongs to the path in VRNN / SRNN -- it is called inside compute_elbo -- but to the
        # model body in Clockwork-VAE, whose forward builds the masks itself, clockwork_vae.py:231-240)
        return [f for f in found if "benchmarking-lvms_b200" in f[0] or f[0].endswith(("evaluation/metrics.py", "utils/log_likelihoods.py", "utils/variational.py"))
                or (name != "cwvae" and f[0].endswith("utils/operations.py") and 90 <= f[1] <= 119)]
    print(f"\n[{name}] host syncs per step: reference {len(syncs_ref)} (path: {len(ours(syncs_ref))}), patched {len(syncs)} "
          f"(path: {len(ours(syncs))}); blvm launches {launches}\n   reference: {sorted(set(syncs_ref))}\n   patched:   {sorted(set(syncs))}")
    assert eager["n"] == 0, "the elementwise KL was materialised"
    assert launches == 4, launches          # sample+mode, likelihood, KL (all levels in one launch), finalize
    assert len(ours(syncs)) == 1, syncs     # the one batched read of the metrics
    assert len(ours(syncs_ref)) >= 5


def test_lazy_samples_make_the_patched_step_three_launches(ref):
    """patch_blvm(lazy_samples=True): sample() / mode() of the patched likelihood return promises, so a VRNN training step is three
    launches of ours (likelihood, KL, finalize); the fused sample + mode kernel runs once, when a reconstruction is first read, and
    the mode it returns is the eager one."""
    import blvm_b200 as B
    M = ref
    name = "vrnn"
    x, x_sl, kwargs = inputs(name)
    state = copy.deepcopy(build(M, name).state_dict())
    try:
        B.patch_blvm(lazy_samples=True)
        model = build(M, name).cuda()
        model.load_state_dict(state)
        assert model.vrnn.likelihood.lazy_samples
        torch.manual_seed(5)
        B.reset_launch_count()
        loss, metrics, out = model(x.cuda(), x_sl, **kwargs)
        loss.backward()
        assert B.launch_count() == 3, B.launch_count()
        lazy = [v for v in vars(out).values() if isinstance(v, B.variational.LazyResult)]
        assert len(lazy) >= 2 and all(v._blvm_value is None for v in lazy)
        first = lazy[0] + 0                                   # reading one launches the fused kernel for both
        assert B.launch_count() == 4 and torch.isfinite(first).all() and float(first.abs().max()) <= 1.0
        second = lazy[1] + 0
        assert B.launch_count() == 4 and second.shape == first.shape
    finally:
        B.unpatch_blvm()
    # the lazily read mode equals the eager one (same kernel, same parameters)
    lik_lazy = B.DiscretizedLogisticMixtureDense(30, 1, 10, 65536, lazy_samples=True).cuda()
    lik = B.DiscretizedLogisticMixtureDense(30, 1, 10, 65536).cuda()
    lik.load_state_dict(lik_lazy.state_dict())
    h = torch.randn(3, 50, 30, device="cuda")
    with torch.no_grad():
        m_lazy, m = lik_lazy.mode(lik_lazy(h)), lik.mode(lik(h))
    assert isinstance(m_lazy, B.variational.LazyResult) and torch.equal(m_lazy + 0, m)
    assert tuple(lik_lazy.sample(lik_lazy(h)).shape) == (3, 50, 1)


@pytest.mark.parametrize("name", ["srnn", "cwvae", "stcn", "stcn_bottom_up", "wavenet", "lstm"])
def test_lazy_samples_and_fused_head_leave_every_model_step_unchanged(name, ref, monkeypatch):
    """patch_blvm(fuse_linear=True, lazy_samples=True) against plain patch_blvm() on the other six models: the fp16-AMP training step
    runs, the loss and the weight gradients are those of the plain patched step (to the fused head's tolerance where it engages), and
    the reconstructions the model returns are promises nobody had to evaluate."""
    import blvm_b200 as B
    from blvm_b200 import ops
    M = ref
    x, x_sl, kwargs = inputs(name)
    state = copy.deepcopy(build(M, name).state_dict())
    res = {}
    calls = {"head": 0}
    real = ops.fused_linear_elbo_apply
    monkeypatch.setattr(ops, "fused_linear_elbo_apply", lambda *a, **k: (calls.__setitem__("head", calls["head"] + 1), real(*a, **k))[1])
    for tag, opts in (("plain", {}), ("lazy", dict(fuse_linear=True, lazy_samples=True))):
        try:
            B.patch_blvm(**opts)
            model = build(M, name).cuda()
            model.load_state_dict(state)
            B.reset_launch_count()
            loss, grads, values, outputs, scaler = train_step(model, x, x_sl, kwargs, True, torch.cuda.amp.GradScaler)
            res[tag] = (float(loss), grads, B.launch_count(), outputs)
        finally:
            B.unpatch_blvm()
        del scaler, model      # a second live GradScaler would make the next run's loss scale ambiguous (amp.active_grad_scaler)
        import gc
        gc.collect()
    (l0, g0, n0, _), (l1, g1, n1, out1) = res["plain"], res["lazy"]
    assert abs(l1 - l0) <= 2e-3 * abs(l0), (l0, l1)
    assert set(g0) == set(g1)
    worst = 0.0
    for k, g in g0.items():
        scale = float(g.abs().max())
        if scale > 0 and torch.isfinite(g).all():
            worst = max(worst, float((g1[k] - g).abs().max()) / scale)
    assert worst < 5e-2, worst
    lazies = [v for v in vars(out1).values() if isinstance(v, B.variational.LazyResult)] if hasattr(out1, "__dict__") else []
    print(f"\n[{name}] plain patched: loss {l0:.9f}, {n0} launches; lazy + fused head: loss {l1:.9f}, {n1} launches, {len(lazies)} promises, "
          f"head calls {calls['head']}, worst grad diff {worst:.2e}")
    assert n1 <= n0
    assert all(v._blvm_value is None for v in lazies)
    assert calls["head"] == (1 if name in ("srnn", "cwvae", "stcn", "stcn_bottom_up") else 0)   # WaveNet: nansum + per-sample output; LSTM: log_prob


def test_wavenet_nansum_keeps_gradient_of_finite_rows(ref):
    """WaveNet.compute_loss reduces with nansum (wavenet.py:143-145): an utterance whose log-prob is NaN drops out of the
    loss and gets a zero upstream gradient, the finite utterances train on.  The real reference method is the checker."""
    from types import SimpleNamespace

    import blvm_b200 as B
    ref_dist = importlib.import_module("blvm.modules.distributions")
    ref_wavenet = importlib.import_module("blvm.models.wavenet.wavenet")
    K, nb, Bn, T = 10, 65536, 4, 700
    g = torch.Generator().manual_seed(2)
    y = (torch.randint(0, nb, (Bn, T, 1), generator=g).float() / (nb - 1) * 2 - 1).cuda()
    raw0 = torch.randn(Bn, T, 3 * K, generator=g)
    raw0[..., K:2 * K] = y.cpu() + 0.05 * raw0[..., K:2 * K]
    raw0[..., 2 * K:] = raw0[..., 2 * K:] * 2 - 4
    raw0[1, 5, K + 3] = float("nan")          # one NaN location inside the valid part of utterance 1
    x_sl = torch.tensor([T, T - 100, T // 2, 9])

    lik_ref = ref_dist.DiscretizedLogisticMixtureDense(x_dim=3 * K, y_dim=1, num_mix=K, num_bins=nb)
    raw_r = raw0.cuda().double().requires_grad_(True)      # the reference's code in fp64: the parity anchor (tests/parity.py)
    lls = raw_r[..., K:].view(Bn, T, 1, 2 * K)
    params_ref = (raw_r[..., :K], lls[..., :K], lls[..., K:].clamp(min=-7.0))          # distributions.py:383-386
    loss_r, logp_r, _ = ref_wavenet.WaveNet.compute_loss(SimpleNamespace(likelihood=lik_ref), y.double(), x_sl, params_ref)
    loss_r.backward()

    raw_o = raw0.cuda().requires_grad_(True)
    lik = B.DiscretizedLogisticMixtureDense(3 * K, 1, K, nb)
    loss_o, logp_o, _ = B.wavenet_compute_loss(SimpleNamespace(likelihood=lik), y, x_sl, B.DMoLParams(raw_o, K, 1, -7.0))
    loss_o.backward()
    assert torch.isnan(logp_r[1]) and torch.isnan(logp_o[1]) and torch.isfinite(loss_o)
    np.testing.assert_allclose(loss_o.item(), loss_r.item(), rtol=1e-6)
    go, gr = raw_o.grad.double().cpu(), raw_r.grad.cpu()
    finite_rows = [0, 2, 3]
    for g0 in range(0, 3 * K, K):      # per parameter group, relative to the group's scale within the sample (tests/parity.py)
        o, r_ = go[finite_rows][..., g0:g0 + K], gr[finite_rows][..., g0:g0 + K]
        tol = 1e-5 * r_.abs() + 1e-5 * r_.abs().amax(-1, keepdim=True) + 1e-6 / float(x_sl.sum())
        assert bool(((o - r_).abs() <= tol).all()), f"group {g0 // K}"
    assert float(go[finite_rows].abs().max()) > 0
    # the NaN utterance: zero gradient wherever the local derivative is finite, in both implementations
    both_finite = torch.isfinite(go[1]) & torch.isfinite(gr[1])
    assert both_finite.float().mean() > 0.99
    assert float(go[1][both_finite].abs().max()) == 0.0 and float(gr[1][both_finite].abs().max()) == 0.0

"""The fused likelihood head (csrc/linear_dmol_kernel.cuh): nn.Linear(x_dim -> 3K) on the tcgen05 tensor
cores, the DMoL value + gradient in registers, and the Linear's backward (dx, dW, db) on the tensor cores, in one kernel.

Reference lines replaced: `DiscretizedLogisticMixtureDense.forward` (blvm/modules/distributions.py:381-387) +
`discretized_logistic_mixture_ll` (blvm/utils/log_likelihoods.py:170-231) + their autograd backward, for 16-bit (AMP) activations.

What is checked, in the order the kernel computes it:
  * RAW = x W^T + b from the tensor cores against an fp32 matmul of the same 16-bit operands (1e-6: fp32 accumulation);
  * the per-sample log-prob and gradient rows are BIT-IDENTICAL to the tile kernel evaluated on that RAW (same device function);
  * dx = G W, dW = G^T x, db = sum G (G rounded to the activation dtype, as an unfused AMP backward would see it) against
    fp32 matmuls, to the activation dtype's resolution / 1e-4 of the tensor scale;
  * through the public API (module with fuse_linear=True -> fused_elbo -> backward, bf16 and fp16 + GradScaler) against the
    unfused path evaluated on fp32 parameters built from the same 16-bit operands; the fused path is the more accurate of the two
    (the unfused AMP path rounds RAW to 16 bits before the likelihood sees it).
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

K, NB = 10, 65536


@pytest.fixture(scope="module")
def B():
    import blvm_b200
    return blvm_b200


def make(Bn, T, Din, dt, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = (torch.randint(0, NB, (Bn, T), device="cuda", generator=g).float() / (NB - 1) * 2 - 1)
    x = torch.randn(Bn, T, Din, device="cuda", generator=g).to(dt)
    W = (torch.randn(3 * K, Din, device="cuda", generator=g) * 0.3)
    b = torch.randn(3 * K, device="cuda", generator=g) * 0.5
    b[2 * K:] -= 4.0
    x_sl = torch.tensor([T] + [max(1, T - 37 * i) for i in range(1, Bn)], dtype=torch.int64)
    return y, x, W, b, x_sl


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 128, 30), (3, 1000, 30), (2, 777, 64), (4, 4096, 32), (5, 301, 30), (2, 130, 78), (1, 1, 2)])
def test_kernel_against_matmuls_and_tile_kernel(B, dt, shape):
    from blvm_b200 import ops
    from blvm_b200._lib import check, lib
    Bn, T, Din = shape
    y, x, W, b, x_sl = make(Bn, T, Din, dt, seed=T)
    W16 = W.to(dt)
    x_dev = x_sl.cuda()
    denom = float(x_sl.sum())
    code = 2 if dt == torch.bfloat16 else 1
    P = 3 * K
    DP = lib.blvm_linear_dmol_padded_dim(K, Din)
    assert DP > 0
    max_ctas = lib.blvm_linear_dmol_max_ctas()
    lp = torch.empty(Bn, T, device="cuda")
    dx = torch.empty_like(x)
    dwp = torch.full((max_ctas, 32, DP), float("nan"), device="cuda")
    chunks = (T + 127) // 128
    part = torch.empty(Bn, chunks, dtype=torch.float64, device="cuda")
    rawdbg = torch.zeros(Bn * T, 32, device="cuda")
    used = ctypes.c_int64(0)
    gscale = -1.0 / denom * (1024.0 if dt == torch.float16 else 1.0)     # fp16: a loss-scale-sized factor keeps G representable
    check(lib.blvm_linear_dmol_fwd_grad(y.data_ptr(), x.data_ptr(), W16.data_ptr(), b.data_ptr(), code, x_dev.data_ptr(), gscale, None, Bn, T, Din,
                                        K, NB, -7.0, 1, lp.data_ptr(), dx.data_ptr(), dwp.data_ptr(), max_ctas, part.data_ptr(), None,
                                        rawdbg.data_ptr(), ctypes.byref(used), ops._stream()), "blvm_linear_dmol_fwd_grad")
    dW = torch.empty(P, Din, device="cuda")
    db = torch.empty(P, device="cuda")
    check(lib.blvm_linear_dmol_reduce_dw(dwp.data_ptr(), used.value, Din, K, dW.data_ptr(), db.data_ptr(), ops._stream()), "reduce")
    torch.cuda.synchronize()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        # 1. the tensor-core product (the bias enters the GEMM as a 16-bit column)
        raw_ref = x.float().reshape(-1, Din) @ W16.float().t() + b.to(dt).float()
        raw = rawdbg[:, :P]
        assert float((raw - raw_ref).abs().max()) <= 2e-6 * float(raw_ref.abs().max()) + 1e-6
        assert float(rawdbg[:, P:].abs().max()) == 0.0
        # 2. value and gradient rows: the same device function as the tile kernel, evaluated on the same fp32 RAW -> bit-identical
        raw_t = raw.reshape(Bn, T, P).contiguous()
        lp2 = torch.empty(Bn, T, device="cuda")
        graw = torch.empty_like(raw_t)
        part2 = torch.empty(Bn * int(lib.blvm_dmol_chunks(T, K, 1)), dtype=torch.float64, device="cuda")
        ops._dmol_call(y, raw_t, x_dev, None, gscale, Bn, T, K, 1, NB, -7.0, 1, lp2, graw, part2)
        torch.cuda.synchronize()
        assert torch.equal(lp, lp2)
        np.testing.assert_allclose(part.sum(1).cpu().numpy(), part2.view(Bn, -1).sum(1).cpu().numpy(), rtol=1e-12, atol=1e-9)
        # 3. the Linear's backward, from G rounded to the activation dtype
        g16 = graw.to(dt).float().reshape(-1, P)
        dx_ref = g16 @ W16.float()
        eps = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
        tol = eps * dx_ref.abs() + 1e-5 * float(dx_ref.abs().max()) + (6e-8 if dt == torch.float16 else 0.0)
        assert bool(((dx.float().reshape(-1, Din) - dx_ref).abs() <= tol).all())
        dW_ref = g16.t() @ x.float().reshape(-1, Din)
        assert float((dW - dW_ref).abs().max()) <= 2e-4 * float(dW_ref.abs().max()) + 1e-12
        db_ref = g16.sum(0)
        assert float((db - db_ref).abs().max()) <= 2e-4 * float(db_ref.abs().max()) + 1e-12
        # padded samples: exact zeros in dx
        for i, n in enumerate(x_sl.tolist()):
            assert float(dx[i, n:].float().abs().max() if n < T else 0.0) == 0.0
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_public_api_fused_head_matches_unfused_fp32_evaluation(B, dt):
    """likelihood(h) under autocast with fuse_linear=True -> fused_elbo -> backward: one tensor-core kernel for the head."""
    from blvm_b200 import ops
    Bn, T, Din, S, Z = 4, 2000, 30, 64, 16
    y, x, W, b, x_sl = make(Bn, T, Din, dt, seed=5)
    lik = B.DiscretizedLogisticMixtureDense(Din, 1, K, NB, fuse_linear=True).cuda()
    with torch.no_grad():
        lik.params.weight.copy_(W)
        lik.params.bias.copy_(b)
    g = torch.Generator(device="cuda").manual_seed(9)
    Tz = -(-T // S)
    kl = [torch.randn(Bn, Tz, Z, device="cuda", generator=g), torch.rand(Bn, Tz, Z, device="cuda", generator=g) + 0.2,
          torch.randn(Bn, Tz, Z, device="cuda", generator=g), torch.rand(Bn, Tz, Z, device="cuda", generator=g) + 0.2]
    scaler = torch.amp.GradScaler("cuda", init_scale=4096.0) if dt == torch.float16 else None

    def run(fused):
        h = x.clone().requires_grad_(True)
        kls = [t.clone().requires_grad_(True) for t in kl]
        lik.zero_grad(set_to_none=True)
        ops.reset_launch_count()
        if fused:
            with torch.autocast("cuda", dtype=dt):
                params = lik(h)
            assert isinstance(params, B.LinearDMoLParams) and not params.materialized
            out = B.fused_elbo(y, params, x_sl, [B.KLLevel(*kls, stride=S)], 0.5, 0.25, num_bins=NB, grad_scaler=scaler)
            assert not params.materialized                     # the (B, T, 3K) tensor was never built
        else:   # the same 16-bit operands, RAW formed in fp32 (what the tensor cores accumulate), likelihood kernel on it
            raw = torch.nn.functional.linear(h.float(), lik.params.weight.to(dt).float(), lik.params.bias.to(dt).float())
            out = B.fused_elbo(y, B.DMoLParams(raw, K, 1, -7.0), x_sl, [B.KLLevel(*kls, stride=S)], 0.5, 0.25, num_bins=NB)
        n_launch = ops.launch_count()
        (scaler.scale(out.loss) if scaler is not None else out.loss).backward()
        s = float(scaler.get_scale()) if scaler is not None else 1.0
        return (out, h.grad.float() / s, lik.params.weight.grad.float() / s, lik.params.bias.grad.float() / s, [t.grad / s for t in kls], n_launch)

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        of, dxf, dWf, dbf, gklf, nf = run(True)
        ou, dxu, dWu, dbu, gklu, nu = run(False)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert nf == 4                                         # head kernel, dW reduce, KL (all levels), finalize
    np.testing.assert_allclose(of.loss.item(), ou.loss.item(), rtol=2e-6)
    np.testing.assert_allclose(of.elbo.cpu().numpy(), ou.elbo.cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(of.sums.cpu().numpy(), ou.sums.cpu().numpy(), rtol=2e-6)
    eps = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10   # G and dx are rounded to the activation dtype inside the fused kernel
    assert float((dxf - dxu).abs().max()) <= eps * float(dxu.abs().max())
    # dW / db are formed from G rounded to the activation dtype (as an unfused AMP backward forms them); the comparison partner uses fp32 G
    assert float((dWf - dWu).abs().max()) <= eps * float(dWu.abs().max())
    assert float((dbf - dbu).abs().max()) <= eps * float(dbu.abs().max())
    for a, c in zip(gklf, gklu):
        assert torch.equal(a, c)                           # the KL path is untouched


def test_lazy_parameters_materialise_for_everything_else(B):
    """LinearDMoLParams behaves like the reference's parameter tuple for every other consumer (sample / mode / indexing /
    log_prob): the Linear is then evaluated once (cuBLAS) and the fused path is simply not taken."""
    Bn, T, Din = 2, 300, 30
    y, x, W, b, x_sl = make(Bn, T, Din, torch.bfloat16, seed=3)
    lik = B.DiscretizedLogisticMixtureDense(Din, 1, K, NB, fuse_linear=True).cuda()
    plain = B.DiscretizedLogisticMixtureDense(Din, 1, K, NB).cuda()
    plain.load_state_dict(lik.state_dict())
    with torch.autocast("cuda", dtype=torch.bfloat16):
        p, q = lik(x), plain(x)
        assert isinstance(p, B.LinearDMoLParams) and not isinstance(q, B.LinearDMoLParams)
        m = lik.mode(p)
        assert m.shape == (Bn, T, 1) and not p.materialized        # a promise: nothing evaluated yet
        assert torch.equal(m + 0, plain.mode(q)) and p.materialized   # reading it evaluates the Linear (cuBLAS) and the sample/mode kernel
        assert torch.equal(p.raw, q.raw) and torch.equal(p[2], q[2]) and len(p) == 3
        out = B.fused_elbo(y, p, x_sl, (), num_bins=NB)    # materialised: the ordinary path
        ref = B.fused_elbo(y, q, x_sl, (), num_bins=NB)
    assert out.loss.item() == ref.loss.item()
    # fp32 activations (no autocast): the module returns ordinary packed parameters
    assert not isinstance(lik(x.float()), B.LinearDMoLParams)

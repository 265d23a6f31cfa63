"""Development check of the fused likelihood head (tcgen05): every GEMM against torch, the DMoL part against the tile kernel."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blvm_b200  # noqa: E402
from blvm_b200 import ops  # noqa: E402
from blvm_b200._lib import lib, check  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
dt = torch.bfloat16 if "--fp16" not in sys.argv else torch.float16
code = 2 if dt == torch.bfloat16 else 1
K, nb = 10, 65536
for (B, T, Din) in [(2, 128, 30), (3, 1000, 30), (2, 777, 64), (4, 4096, 32), (256, 16000, 30)]:
    P = 3 * K
    y = (torch.randint(0, nb, (B, T), device=dev).float() / (nb - 1) * 2 - 1)
    x = torch.randn(B, T, Din, device=dev).to(dt)
    W = (torch.randn(P, Din, device=dev) * 0.3).to(dt)
    b = torch.randn(P, device=dev) * 0.5
    b[2 * K:] = b[2 * K:] - 4
    x_sl = torch.tensor([T] + [max(1, T - 37 * i) for i in range(1, B)], dtype=torch.int64, device=dev)
    denom = float(x_sl.sum())
    lp = torch.empty(B, T, device=dev)
    dx = torch.empty_like(x)
    DP = lib.blvm_linear_dmol_padded_dim(K, Din)
    max_ctas = lib.blvm_linear_dmol_max_ctas()
    dwp = torch.zeros(max_ctas, 32, DP, device=dev)
    chunks = (T + 127) // 128
    part = torch.empty(B, chunks, dtype=torch.float64, device=dev)
    rawdbg = torch.zeros(B * T, 32, device=dev)
    used = ctypes.c_int64(0)
    check(lib.blvm_linear_dmol_fwd_grad(y.data_ptr(), x.data_ptr(), W.data_ptr(), b.data_ptr(), code, x_sl.data_ptr(), -1.0 / denom, None,
                                        B, T, Din, K, nb, -7.0, 1, lp.data_ptr(), dx.data_ptr(), dwp.data_ptr(), max_ctas, part.data_ptr(), None,
                                        rawdbg.data_ptr(), ctypes.byref(used), ops._stream()), "linear_dmol")
    dW = torch.empty(P, Din, device=dev)
    db = torch.empty(P, device=dev)
    check(lib.blvm_linear_dmol_reduce_dw(dwp.data_ptr(), used.value, Din, K, dW.data_ptr(), db.data_ptr(), ops._stream()), "reduce")
    torch.cuda.synchronize()
    # references
    raw_ref = x.float().reshape(-1, Din) @ W.float().t() + b.to(dt).float()     # the bias enters the GEMM in the 16-bit dtype
    e_raw = (rawdbg[:, :P] - raw_ref).abs().max().item() / raw_ref.abs().max().item()
    raw_t = raw_ref.reshape(B, T, P).contiguous()
    lp2 = torch.empty(B, T, device=dev)
    graw = torch.empty_like(raw_t)
    part2 = torch.empty(B * int(lib.blvm_dmol_chunks(T, K, 1)), dtype=torch.float64, device=dev)
    ops._dmol_call(y, raw_t, x_sl, None, -1.0 / denom, B, T, K, 1, nb, -7.0, 1, lp2, graw, part2)
    torch.cuda.synchronize()
    e_lp = (lp - lp2).abs().max().item()
    g16 = graw.to(dt).float().reshape(-1, P)
    dx_ref = g16 @ W.float()
    dW_ref = g16.t() @ x.float().reshape(-1, Din)
    db_ref = g16.sum(0)
    e_dx = (dx.float().reshape(-1, Din) - dx_ref).abs().max().item() / dx_ref.abs().max().item()
    e_dW = (dW - dW_ref).abs().max().item() / dW_ref.abs().max().item()
    e_db = (db - db_ref).abs().max().item() / db_ref.abs().max().item()
    e_part = (part.sum(1) - part2.view(B, -1).sum(1)).abs().max().item() / part2.abs().sum().item()
    print(f"B={B} T={T} Din={Din} DP={DP} ctas={used.value}: raw {e_raw:.2e}  lp {e_lp:.2e}  dx {e_dx:.2e}  dW {e_dW:.2e}  db {e_db:.2e}  rowsum {e_part:.2e}")
    if B == 256:
        import time
        for _ in range(3):
            lib.blvm_linear_dmol_fwd_grad(y.data_ptr(), x.data_ptr(), W.data_ptr(), b.data_ptr(), code, x_sl.data_ptr(), -1.0 / denom, None,
                                          B, T, Din, K, nb, -7.0, 1, lp.data_ptr(), dx.data_ptr(), dwp.data_ptr(), max_ctas, part.data_ptr(), None,
                                          None, None, ops._stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            lib.blvm_linear_dmol_fwd_grad(y.data_ptr(), x.data_ptr(), W.data_ptr(), b.data_ptr(), code, x_sl.data_ptr(), -1.0 / denom, None,
                                          B, T, Din, K, nb, -7.0, 1, lp.data_ptr(), dx.data_ptr(), dwp.data_ptr(), max_ctas, part.data_ptr(), None,
                                          None, None, ops._stream())
        e1.record()
        torch.cuda.synchronize()
        print(f"   fused head fwd+bwd: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch ({B * T * 128 / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9:.0f} GB/s of 128 B/sample)")
        # unfused: F.linear fwd, dmol kernel, two backward GEMMs
        xr = x.detach().clone().requires_grad_(True)
        Wr = W.detach().clone().requires_grad_(True)
        br = b.to(dt).detach().clone().requires_grad_(True)
        def unfused():
            raw = torch.nn.functional.linear(xr, Wr, br)
            g = torch.empty_like(raw)
            ops._dmol_call(y, raw.detach(), x_sl, None, -1.0 / denom, B, T, K, 1, nb, -7.0, 1, lp2, g, part2)
            raw.backward(g)
            xr.grad = None; Wr.grad = None; br.grad = None
        for _ in range(3):
            unfused()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            unfused()
        e1.record()
        torch.cuda.synchronize()
        print(f"   unfused (cuBLAS linear + DMoL kernel + cuBLAS backward): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")

"""ctypes access to the C oracle (oracle/blvm_oracle.c).  TEST INFRASTRUCTURE: used by tests/test_oracle_c.py and by
bench.py's cpu_baseline / --impl reference arm only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libblvm_oracle.so")
_lib = None

_f = {np.float32: (ctypes.c_float, "f32"), np.float64: (ctypes.c_double, "f64")}


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_elbo_loss.restype = ctypes.c_double
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(load().oracle_num_threads())


def use_all_cores() -> int:
    """Use every host core regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().oracle_set_num_threads(int(n))
    return num_threads()


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def dmol(y, raw, x_sl, K, num_bins, log_eps=-7.0, gscale=1.0, want_grad=True, want_lp=True):
    """y (B, T), raw (B, T, 3K), x_sl (B) -> masked lp (B, T), graw, row_logp (B) fp64.  dtype of `raw` selects fp32/fp64."""
    lib = load()
    dt = raw.dtype.type
    ct, sfx = _f[dt]
    B, T = y.shape
    y = np.ascontiguousarray(y, dt)
    raw = np.ascontiguousarray(raw, dt)
    x_sl = np.ascontiguousarray(x_sl, np.int64)
    lp = np.empty((B, T), dt) if want_lp else None
    graw = np.empty_like(raw) if want_grad else None
    rows = np.empty(B, np.float64)
    fn = getattr(lib, "oracle_dmol_" + sfx)
    fn(_p(y, ct), _p(raw, ct), _p(x_sl, ctypes.c_int64), ctypes.c_int64(B), ctypes.c_int64(T), int(K), int(num_bins),
       ct(log_eps), ct(gscale), _p(lp, ct), _p(graw, ct), _p(rows, ctypes.c_double))
    return lp, graw, rows


def kl(mu_q, sd_q, mu_p, sd_p, lens, free_nats, gscale=1.0, want_grad=True):
    lib = load()
    dt = mu_q.dtype.type
    ct, sfx = _f[dt]
    B, Tz, Z = mu_q.shape
    ins = [np.ascontiguousarray(a, dt) for a in (mu_q, sd_q, mu_p, sd_p)]
    lens = np.ascontiguousarray(lens, np.int64)
    grads = [np.empty_like(ins[0]) for _ in range(4)] if want_grad else [None] * 4
    row_kl, row_fn = np.empty(B, np.float64), np.empty(B, np.float64)
    fn = getattr(lib, "oracle_kl_" + sfx)
    fn(*[_p(a, ct) for a in ins], _p(lens, ctypes.c_int64), ctypes.c_int64(B), ctypes.c_int64(Tz), ctypes.c_int64(Z),
       ctypes.c_double(free_nats), ct(gscale), *[_p(g, ct) for g in grads], _p(row_kl, ctypes.c_double),
       _p(row_fn, ctypes.c_double))
    return grads, row_kl, row_fn


def elbo_loss(row_logp, row_kl, row_klfn, x_sl, beta):
    lib = load()
    B = len(row_logp)
    elbo = np.empty(B, np.float64)
    x_sl = np.ascontiguousarray(x_sl, np.int64)
    d = ctypes.c_double
    loss = lib.oracle_elbo_loss(_p(np.ascontiguousarray(row_logp), d), _p(None if row_kl is None else np.ascontiguousarray(row_kl), d),
                                _p(None if row_klfn is None else np.ascontiguousarray(row_klfn), d), _p(x_sl, ctypes.c_int64),
                                ctypes.c_int64(B), d(beta), _p(elbo, d))
    return float(loss), elbo


def elbo_step(y, raw, x_sl, kl_levels, beta, K, num_bins, want_grad=True):
    """One full step of the path (what bench.py times on the host): DMoL value+grad, KL value+grad per level, loss."""
    denom = float(np.sum(x_sl))
    dt = raw.dtype.type
    _, graw, row_logp = dmol(y, raw, x_sl, K, num_bins, gscale=dt(-1.0 / denom), want_grad=want_grad, want_lp=False)
    row_kl = np.zeros(len(x_sl))
    row_fn = np.zeros(len(x_sl))
    gkl = []
    for lv in kl_levels:
        lens = -(-np.asarray(x_sl) // lv["stride"])
        g, a, f = kl(lv["mu_q"], lv["sd_q"], lv["mu_p"], lv["sd_p"], lens, lv["free_nats"], gscale=dt(beta / denom),
                     want_grad=want_grad)
        row_kl += a
        row_fn += f
        gkl.append(g)
    loss, elbo = elbo_loss(row_logp, row_kl, row_fn, x_sl, beta)
    return dict(loss=loss, elbo=elbo, logp=row_logp, kl=row_kl, kl_fn=row_fn, graw=graw, gkl=gkl)

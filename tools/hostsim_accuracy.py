"""Development aid (CPU): worst error / tolerance of the host build of csrc/blvm_math.cuh (MUFU-degraded transcendentals)
against the fp64 oracle, over the goldens and the extreme-regime sweep of tests/test_hostsim_math.py.  Used to compare
formula variants before a GPU run:  python tools/hostsim_accuracy.py [-DFLAG=...]"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from oracle import blvm_oracle as O  # noqa: E402

FP = ctypes.POINTER(ctypes.c_float)
P = lambda a: a.ctypes.data_as(FP)  # noqa: E731


def build(extra):
    out = os.path.join(ROOT, "tests", "hostsim", "_build", "libhostsim_probe.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-DBLVM_HOSTSIM_MUFU_BITS=2", *extra, "-I",
                    os.path.join(ROOT, "benchmarking-lvms_b200", "csrc"), "-o", out, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")], check=True)
    return ctypes.CDLL(out)


def worst(lp, gr, L, G, K, gout, ok=None):
    e_lp = np.abs(lp - L) / (1e-5 * np.abs(L) + 1e-6)
    w = 0.0
    gout = np.abs(gout).reshape(-1, 1)
    for g0 in range(0, gr.shape[1], K):
        o, r = gr[:, g0:g0 + K].astype(np.float64), G[:, g0:g0 + K]
        gmax = np.abs(r).max(-1, keepdims=True)
        tol = 1e-5 * np.abs(r) + 1e-5 * gmax + 1e-6 * gout + 1e-30
        ratio = np.abs(o - r) / tol
        if ok is not None:
            ratio = ratio[ok]
        w = max(w, ratio.max())
    if ok is not None:
        e_lp = e_lp[ok]
    return e_lp.max(), w


def main():
    sim = build(sys.argv[1:])
    for case in ["dmol_K1_nb65536", "dmol_K2_nb65536", "dmol_K10_nb65536", "dmol_K30_nb65536", "dmol_K10_nb256"]:
        g = load_golden(case)
        K, nb = int(g["K"]), int(g["num_bins"])
        y, raw, gout = (np.ascontiguousarray(g[k], np.float32) for k in ("y", "raw", "gout"))
        N = raw.shape[0]
        lp, gr = np.empty(N, np.float32), np.empty_like(raw)
        sim.hostsim_dmol(P(y), P(raw), P(gout), ctypes.c_int64(N), K, 1, nb, ctypes.c_float(-7.0), 0, P(lp), P(gr))
        a, b = worst(lp.astype(np.float64), gr, g["lp64"].reshape(-1), g["graw64"].reshape(N, -1), K, gout)
        print(f"{case:22s} lp err/tol {a:.3f}  grad err/tol {b:.3f}")
    for K in (2, 10, 30):
        nb = 65536
        rng = np.random.default_rng(1000 + K + nb)
        N = 20000
        y = (rng.integers(0, nb, N) / (nb - 1) * 2 - 1).astype(np.float32)
        y[:40] = rng.choice(np.array([-1.0, 1.0, 2 / nb - 1, 1 - 2 / nb], np.float32), 40)
        raw = np.empty((N, 3 * K), np.float32)
        raw[:, :K] = rng.normal(0, 1, (N, K)) * rng.choice([1, 10, 40], (N, 1))
        raw[:, K:2 * K] = y[:, None] + rng.normal(0, 1, (N, K)) * rng.choice([1e-4, 1e-2, 0.3, 3.0], (N, K))
        raw[:, 2 * K:] = rng.uniform(-12, 4, (N, K))
        gout = rng.normal(0, 1, N).astype(np.float32)
        lp, gr = np.empty(N, np.float32), np.empty_like(raw)
        sim.hostsim_dmol(P(y), P(raw), P(gout), ctypes.c_int64(N), K, 1, nb, ctypes.c_float(-7.0), 0, P(lp), P(gr))
        L, G = O.dmol_value_and_grad(y.astype(np.float64), raw.astype(np.float64), K, 1, nb, -7.0, gout.astype(np.float64))
        _, delta = O.dmol_branches(y.astype(np.float64).reshape(N, 1), raw.astype(np.float64), K, 1, nb, -7.0)
        ok = ~(np.abs(delta / 1e-5 - 1) < 2e-4).any(axis=(1, 2))
        a, b = worst(lp.astype(np.float64), gr, L, G, K, gout, ok)
        print(f"extreme K={K:2d} nb={nb}   lp err/tol {a:.3f}  grad err/tol {b:.3f}")


if __name__ == "__main__":
    main()

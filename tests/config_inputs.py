"""Seeded synthetic inputs for BASELINE.json's configs 1-4 at FULL size (shared by tests/golden/make_golden_configs.py,
which runs the reference on them in the build container, and tests/test_gpu_configs.py, which runs the CUDA path on the
same arrays on the GPU box).  numpy's default_rng is bit-reproducible across machines, so only the reference OUTPUTS are
stored in tests/golden/configs_full.npz."""
import numpy as np

NUM_BINS = 65536
K = 10

CONFIGS = {
    # name: model reducer, B, T, latent levels [(overall stride, Z)], beta, free_nats
    "config1_vrnn": dict(model="vrnn", B=4, T=16000, levels=[(200, 64)], beta=0.5, free_nats=0.0625, seed=11),
    "config2_wavenet": dict(model="wavenet", B=32, T=16000, levels=[], beta=1.0, free_nats=0.0, seed=12),
    "config3_srnn": dict(model="srnn", B=64, T=32000, levels=[(64, 64)], beta=0.5, free_nats=0.0625, seed=13),
    "config4_cwvae": dict(model="cwvae", B=32, T=65536, levels=[(64, 128), (512, 64), (4096, 32)], beta=0.5,
                          free_nats=0.0625, seed=14),
}


def make_inputs(name):
    c = CONFIGS[name]
    rng = np.random.default_rng(c["seed"])
    B, T = c["B"], c["T"]
    y = (rng.integers(0, NUM_BINS, (B, T)).astype(np.float32) / np.float32(NUM_BINS - 1) * 2 - 1).astype(np.float32)
    raw = rng.standard_normal((B, T, 3 * K), dtype=np.float32)
    raw[..., K:2 * K] = y[..., None] + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    x_sl = (T * rng.uniform(0.5, 1.0, B)).astype(np.int64)
    x_sl[0] = T                                   # the batch is padded to its longest utterance
    kl = []
    for stride, Z in c["levels"]:
        Tz = -(-T // stride)
        lv = [rng.standard_normal((B, Tz, Z), dtype=np.float32) for _ in range(4)]
        for i in (1, 3):
            lv[i] = (np.log1p(np.exp(lv[i])) + 1e-3).astype(np.float32)
        kl.append(lv)                              # [mu_q, sd_q, mu_p, sd_p]
    return y, raw, x_sl, kl


def probe_indices(name, n=256):
    """Fixed random positions at which the reference gradients are stored (the full gradient is hundreds of MB)."""
    c = CONFIGS[name]
    rng = np.random.default_rng(1000 + c["seed"])
    return rng.integers(0, c["B"], n), rng.integers(0, c["T"], n), rng.integers(0, 3 * K, n)

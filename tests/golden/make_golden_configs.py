#!/usr/bin/env python
"""Run the REFERENCE's compute_elbo / compute_loss on BASELINE.json's configs 1-4 at full size (build container only)
and store the outputs + gradient probes in tests/golden/configs_full.npz.  Inputs come from tests/config_inputs.py.

    python tests/golden/make_golden_configs.py
"""
import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

os.environ.setdefault("BLVM_DATA_ROOT_DIRECTORY", "/tmp/blvmdata")
os.makedirs(os.environ["BLVM_DATA_ROOT_DIRECTORY"], exist_ok=True)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref_shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))

from config_inputs import CONFIGS, K, NUM_BINS, make_inputs, probe_indices  # noqa: E402
from blvm.modules.distributions import DiscretizedLogisticMixtureDense  # noqa: E402
from blvm.utils.operations import sequence_mask  # noqa: E402
from blvm.utils.variational import kl_divergence_gaussian  # noqa: E402
import blvm.models  # noqa: E402,F401

ref = {m: importlib.import_module(p) for m, p in dict(vrnn="blvm.models.vrnn", srnn="blvm.models.srnn",
                                                       cwvae="blvm.models.clockwork_vae.clockwork_vae",
                                                       wavenet="blvm.models.wavenet.wavenet").items()}


def params_from_raw(raw):
    logit_probs = raw[..., :K]
    lls = raw[..., K:].view(*raw.shape[:-1], 1, 2 * K)
    locs, log_scales = lls.chunk(2, dim=-1)
    return logit_probs, locs, log_scales.clamp(min=-7.0)     # distributions.py:383-387


def run(name, dt):
    c = CONFIGS[name]
    y, raw, x_sl, kl = make_inputs(name)
    y_t = torch.from_numpy(y).to(dt).unsqueeze(-1)
    raw_t = torch.from_numpy(raw).to(dt).requires_grad_(True)
    x_sl_t = torch.from_numpy(x_sl)
    kl_t = [[torch.from_numpy(a).to(dt).requires_grad_(True) for a in lv] for lv in kl]
    lik = DiscretizedLogisticMixtureDense(x_dim=3 * K, y_dim=1, num_mix=K, num_bins=NUM_BINS)
    T = c["T"]
    if c["model"] in ("vrnn", "srnn"):
        cls = ref["vrnn"].VRNN if c["model"] == "vrnn" else ref["srnn"].SRNN
        kld = kl_divergence_gaussian(*kl_t[0])
        loss, elbo, logp, klr, _ = cls.compute_elbo(SimpleNamespace(likelihood=lik), y_t, params_from_raw(raw_t), kld, x_sl_t,
                                                    c["levels"][0][0], c["beta"], c["free_nats"])
    elif c["model"] == "cwvae":
        ostr = [s for s, _ in c["levels"]]
        self = SimpleNamespace(likelihood=lik, num_levels=len(ostr), overall_strides=ostr)
        seq_mask = sequence_mask(x_sl_t, max_len=T)
        level_masks = [sequence_mask((x_sl_t / s).ceil().int(), max_len=-(-T // s)) for s in ostr]
        klds = [kl_divergence_gaussian(*lv) for lv in kl_t]
        loss, elbo, logp, klr, _ = ref["cwvae"].CWVAE.compute_elbo(self, y_t, seq_mask, level_masks, x_sl_t,
                                                                   params_from_raw(raw_t), klds, c["beta"], c["free_nats"])
    else:
        loss, logp, _ = ref["wavenet"].WaveNet.compute_loss(SimpleNamespace(likelihood=lik), y_t, x_sl_t, params_from_raw(raw_t))
        elbo, klr = logp, torch.zeros_like(logp)
    loss.backward()
    bi, ti, pi = probe_indices(name)
    grp = (pi // K) * K                                        # parameter group (logits / locs / log-scales) of each probe
    scale = torch.stack([raw_t.grad[bi, ti, g0:g0 + K].abs().max() for bi, ti, g0 in zip(bi.tolist(), ti.tolist(), grp.tolist())])
    out = dict(loss=loss.item(), graw_probe_scale=scale.double().numpy(), elbo=elbo.detach().double().numpy(), logp=logp.detach().double().numpy(),
               kl=klr.detach().double().numpy(), graw_probe=raw_t.grad[bi, ti, pi].double().numpy(),
               graw_abs_sum=raw_t.grad.double().abs().sum().item(), graw_sum=raw_t.grad.double().sum().item())
    for l, lv in enumerate(kl_t):
        for nme, t in zip(("mu_q", "sd_q", "mu_p", "sd_p"), lv):
            out[f"g_{nme}_{l}_abs_sum"] = t.grad.double().abs().sum().item()
            out[f"g_{nme}_{l}_probe"] = t.grad.reshape(-1)[:: max(1, t.grad.numel() // 256)][:256].double().numpy()
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    store = {}
    for name in CONFIGS:
        for tag, dt in (("64", torch.float64), ("32", torch.float32)):
            r = run(name, dt)
            for k, v in r.items():
                store[f"{name}/{k}{tag}"] = np.asarray(v)
            print(name, tag, "loss", r["loss"])
    path = os.path.join(HERE, "configs_full.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")

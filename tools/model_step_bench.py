"""Model-level effect of the drop-in: the reference's OWN audio models (staged, unmodified, oracle/_ref) at the shapes of BASELINE configs
1-4 with the experiment scripts' default sizes, one AMP training step (experiments/experiment_vrnn_audio.py:216-232: fp16 autocast,
GradScaler, backward, unscale, clip, step) timed end to end on the GPU -- unpatched, under patch_blvm(), and under
patch_blvm(fuse_linear=True, lazy_samples=True).  Wall clock with a device synchronisation on both sides (host-inclusive, eager).

    python tools/model_step_bench.py [config1_vrnn config2_wavenet config3_srnn config4_cwvae] > profiles/r2_model_steps.jsonl
"""
import gc
import importlib
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

ref_loader.load()
import blvm.models as M  # noqa: E402

import blvm_b200 as B  # noqa: E402

NB = 2 ** 16


def build(name):
    torch.manual_seed(3)
    if name == "config1_vrnn":        # experiment_vrnn_audio.py:171-181 defaults
        return M.VRNNAudio(input_size=200, hidden_size=512, latent_size=256, likelihood="DMoL", num_bins=NB), 4, 16000, dict(beta=1.0, free_nats=0.0625)
    if name == "config2_wavenet":     # experiment_wavenet_audio.py:152-174 defaults
        lik = importlib.import_module("blvm.modules.distributions").DiscretizedLogisticMixtureDense
        return M.WaveNet(likelihood=lik(x_dim=64, y_dim=1, num_mix=10, num_bins=NB), n_layers=10, n_stacks=4, res_channels=64, num_bins=NB), 32, 16000, {}
    if name == "config3_srnn":        # experiment_srnn_audio.py:172-182, stack 64 (BASELINE config 3), z 256 (benchmarks.txt:22)
        return M.SRNNAudio(likelihood="DMoL", input_size=64, hidden_size=512, latent_size=256, num_bins=NB), 64, 32000, dict(beta=1.0, free_nats=0.0625)
    if name == "config4_cwvae":       # experiment_clockwork_audio.py:84-96 with strides 64/8/8 (BASELINE config 4)
        # batch 8 instead of 32: with the default hidden size 512 the reference itself needs > 170 GB at 32 x 65536 samples
        return M.CWVAEAudio(z_size=[128, 64, 32], h_size=[512, 512, 512], strides=[64, 8, 8], num_level_layers=8, stride_per_layer=2,
                            likelihood="DMoL", num_bins=NB), 8, 65536, dict(beta=1.0, free_nats=4.0)
    raise KeyError(name)


def step(model, opt, scaler, x, x_sl, kw):
    with torch.autocast("cuda", dtype=torch.float16):
        loss, metrics, outputs = model(x, x_sl, **kw)
    opt.zero_grad(set_to_none=True)
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    torch.nn.utils.clip_grad_value_(model.parameters(), 3000.0)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 3000.0)
    scaler.step(opt)
    scaler.update()
    vals = [float(m.value) for m in metrics]      # the training loop reads its metrics every step
    return float(loss.detach()), vals


WARMUP = 3   # TorchScript (VRNNCell / RSSMCell are scripted) profiles and optimises during the first calls


def run(name, variant, iters):
    patched = not variant.startswith("reference")
    if patched:
        B.patch_blvm(fuse_linear=(variant == "patched+head+lazy"), lazy_samples=(variant == "patched+head+lazy"))
    try:
        model, Bn, T, kw = build(name)
        model = model.cuda().train()
        g = torch.Generator().manual_seed(1)
        x = (torch.randint(0, NB, (Bn, T), generator=g).float() / (NB - 1) * 2 - 1).cuda()
        x_sl = torch.full((Bn,), T, dtype=torch.int64)
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        scaler = torch.amp.GradScaler("cuda")
        torch.cuda.reset_peak_memory_stats()
        times, loss = [], None
        for i in range(iters + WARMUP):
            torch.manual_seed(5)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loss, _ = step(model, opt, scaler, x, x_sl, kw)
            torch.cuda.synchronize()
            if i >= WARMUP:
                times.append(time.perf_counter() - t0)
        return {"config": name, "variant": variant, "B": Bn, "T": T, "ms_per_step": statistics.median(times) * 1e3, "steps_timed": iters,
                "samples_per_s": Bn * T / statistics.median(times), "loss": loss, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                "blvm_launches_per_step": (B.launch_count() // (iters + WARMUP)) if patched else 0}
    finally:
        if patched:
            B.unpatch_blvm()
        B.reset_launch_count()
        gc.collect()
        torch.cuda.empty_cache()


def main():
    names = sys.argv[1:] or ["config1_vrnn", "config2_wavenet", "config3_srnn", "config4_cwvae"]
    for name in names:
        base = None
        for variant in ("reference", "patched", "patched+head+lazy", "reference (again)"):
            try:
                r = run(name, variant, 5)
                if variant == "reference":
                    base = r["ms_per_step"]
                r["speedup_vs_reference"] = (base / r["ms_per_step"]) if base else None
            except Exception as err:   # one model must not take the table down
                r = {"config": name, "variant": variant, "error": repr(err)[:300]}
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()

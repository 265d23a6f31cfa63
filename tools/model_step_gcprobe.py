"""Development aid: is the patched reference-model step slower on the host because of Python's cyclic GC?"""
import gc
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import model_step_bench as msb  # noqa: E402

B = msb.B
name = sys.argv[1] if len(sys.argv) > 1 else "config3_srnn"
stats = {"n": 0, "t": 0.0, "t0": 0.0}


def cb(phase, info):
    if phase == "start":
        stats["t0"] = time.perf_counter()
    else:
        stats["n"] += 1
        stats["t"] += time.perf_counter() - stats["t0"]


gc.callbacks.append(cb)
for gc_on in (True, False):
    for variant in ("reference", "patched"):
        if variant == "patched":
            B.patch_blvm()
        model, Bn, T, kw = msb.build(name)
        model = model.cuda().train()
        g = torch.Generator().manual_seed(1)
        x = (torch.randint(0, msb.NB, (Bn, T), generator=g).float() / (msb.NB - 1) * 2 - 1).cuda()
        x_sl = torch.full((Bn,), T, dtype=torch.int64)
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        scaler = torch.amp.GradScaler("cuda")
        for _ in range(4):
            msb.step(model, opt, scaler, x, x_sl, kw)
        gc.collect()
        (gc.enable if gc_on else gc.disable)()
        stats.update(n=0, t=0.0)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            msb.step(model, opt, scaler, x, x_sl, kw)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        gc.enable()
        print(f"{name} {variant:10s} gc={'on ' if gc_on else 'off'}: {statistics.median(ts) * 1e3:7.1f} ms/step; gc runs {stats['n']} taking {stats['t'] * 1e3:.1f} ms over 5 steps")
        if variant == "patched":
            B.unpatch_blvm()
        del model, opt, scaler
        gc.collect()
        torch.cuda.empty_cache()

"""BASELINE config 5 sweep: DMoL+KL ELBO step at B=256, T in {16000..128000}, K in {1,10,30} on one GPU.
Prints one markdown row per point (kernel-only roofline of the DMoL kernel and whole-step throughput)."""
import json
import subprocess
import sys

rows = []
for K in (1, 10, 30):
    for T in (16000, 32000, 64000, 128000):
        steps = max(20, int(400 * 16000 / T * 10 / max(K, 10)))
        out = subprocess.run([sys.executable, "bench.py", "--K", str(K), "--T", str(T), "--steps", str(steps), "--warmup", "5",
                              "--no-e2e", "--no-cpu-baseline", "--min-seconds", "1.0"], capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            print(f"| {K} | {T} | failed: {out.stderr[-200:]!r} |")
            continue
        r = d["roofline"]
        print(f"| {K} | {T} | {r['us_per_launch']:.1f} | {r['achieved']:.0f} | {r['frac']:.3f} | {r['frac_of_nominal_8TBs']:.3f} | "
              f"{d['ms_per_step']:.4f} | {d['value'] / 1e9:.2f} | {d['clocks']['sm_mhz']} |", flush=True)

"""A/B harness for dmol_stream_kernel tuning: builds variant libraries (stages / lookahead / CTA size / K cap) in
parallel with nvcc and times each against the tile kernel on the GPU box.

    python tools/ab_stream.py build          # here (no GPU): benchmarking-lvms_b200/lib/variants/*.so
    python tools/ab_stream.py run [--Ks 1 5] # on the GPU box (gpurun): one line per (variant, K, mode)
"""
import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "benchmarking-lvms_b200")
VAR_DIR = os.path.join(PKG, "lib", "variants")
# tag: (stages, lookahead, tpb[, samples per thread for K <= 2, for K <= 5])
VARIANTS = {"s3l1": (3, 1, 128), "s2l1": (2, 1, 128), "s4l2": (4, 2, 128), "s3l2": (3, 2, 128), "s3l1t256": (3, 1, 256), "s4l2t256": (4, 2, 256),
            "s2l1_spt84": (2, 1, 128, 8, 4), "s2l1_spt42": (2, 1, 128, 4, 2), "s2l1_spt44": (2, 1, 128, 4, 4), "s2l1_spt82": (2, 1, 128, 8, 2),
            "s3l1_spt42": (3, 1, 128, 4, 2), "s2l1_spt21": (2, 1, 128, 2, 1)}


def build():
    os.makedirs(VAR_DIR, exist_ok=True)

    def one(tag):
        s, la, tpb = VARIANTS[tag][:3]
        spt = VARIANTS[tag][3:]
        out = os.path.join(VAR_DIR, f"libblvm_b200_{tag}.so")
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
               f"-DBLVM_STREAM_STAGES={s}", f"-DBLVM_STREAM_LOOKAHEAD={la}", f"-DBLVM_STREAM_TPB={tpb}", "-DBLVM_STREAM_MAX_K=10",
               *([f"-DBLVM_SPT_1={spt[0]}", f"-DBLVM_SPT_2={spt[0]}", f"-DBLVM_SPT_5={spt[1]}", "-DBLVM_SPT_8=2", "-DBLVM_SPT_12=1"] if spt else []),
               "-o", out] + [os.path.join(PKG, "csrc", f) for f in ("blvm_b200.cu", "blvm_dmol_f32.cu", "blvm_dmol_f16.cu", "blvm_dmol_bf16.cu")]
        subprocess.run(cmd, check=True)
        return out

    want = [t for t in sys.argv[2:] if t in VARIANTS] or list(VARIANTS)
    tags = [t for t in want if not os.path.exists(os.path.join(VAR_DIR, f"libblvm_b200_{t}.so")) or "--force" in sys.argv]
    with ThreadPoolExecutor(max_workers=6) as pool:
        for o in pool.map(one, tags):
            print("built", o)


CHILD = r"""
import sys, torch
sys.path.insert(0, %(root)r)
import blvm_b200
from blvm_b200 import ops
lib = blvm_b200._lib.lib
B, nb = 256, 65536
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]
for T in %(Ts)r:
  for K in %(Ks)r:
    y = torch.randint(0, nb, (B, T), device='cuda').float() / (nb - 1) * 2 - 1
    raw = torch.randn(B, T, 3 * K, device='cuda')
    raw[..., K:2 * K] = y.unsqueeze(-1) + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    x_dev = torch.full((B,), T, dtype=torch.int64, device='cuda')
    lp = torch.empty(B, T, device='cuda'); graw = torch.empty_like(raw)
    part = torch.empty(B * int(lib.blvm_dmol_chunks(T, K, 1)), dtype=torch.float64, device='cuda')
    for mode in %(modes)r:
        lib.blvm_set_stream_mode(mode)
        t = timeit(lambda: ops._dmol_call(y, raw, x_dev, None, -1e-6, B, T, K, 1, nb, -7.0, 1, lp, graw, part))
        tf = timeit(lambda: ops._dmol_call(y, raw, x_dev, None, 0.0, B, T, K, 1, nb, -7.0, 1, lp, None, part))
        N = B * T
        print(f"%(tag)s T={T} K={K:2d} {'stream' if mode else 'tile  '} fwd+grad {t*1e3:7.1f} us {N*4*(2+6*K)/t/1e6:6.0f} GB/s | fwd {tf*1e3:7.1f} us {N*4*(2+3*K)/tf/1e6:6.0f} GB/s", flush=True)
    del raw, graw
"""


def run(Ks, Ts, tags):
    for i, tag in enumerate(tags):
        env = dict(os.environ, BLVM_B200_LIB=os.path.join(VAR_DIR, f"libblvm_b200_{tag}.so"))
        code = CHILD % dict(root=ROOT, Ks=Ks, Ts=Ts, tag=tag, modes=[0, 1] if i == 0 else [1])
        subprocess.run([sys.executable, "-c", code], env=env, check=False)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["build", "run"])
    ap.add_argument("build_tags", nargs="*")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--Ks", type=int, nargs="+", default=[1, 2, 5, 10])
    ap.add_argument("--Ts", type=int, nargs="+", default=[16000, 64000])
    ap.add_argument("--tags", nargs="+", default=list(VARIANTS))
    a = ap.parse_args()
    build() if a.what == "build" else run(a.Ks, a.Ts, a.tags)

#!/usr/bin/env python
"""Contract benchmark: waveform samples/s of one DMoL+KL ELBO forward+backward step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--T 16000] [--K 10] [--ragged]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload = BASELINE.json configs[4] ("standalone DMoL+KL ELBO kernel sweep", the configuration the metric
"DMoL+KL ELBO fwd+bwd at 1/2/4/8 B200; % HBM roofline" is quoted on) at its first grid point: per GPU 256 utterances x
16000 samples, DMoL K=10, 16-bit bins, one latent layer of stride 64 / width 64 (SRNN-like), beta 0.5, free nats 1/16.
Weak scaling: every rank owns its own 256 utterances; the only exchange is that of the scalar sums (fused into the
finalize kernel over NVLink peer memory, or an NCCL all-reduce).  `--workload config2|config3|config3z256|config4` run the
same step at the shapes of BASELINE configs 2-4 (informational); `--impl reference --reference-device cuda|cpu-torch`
time the reference's own op chain as eager PyTorch on this GPU / on the host threads (informational).

One JSON line on stdout (rank 0).  `value` times K steps with inputs resident in HBM (CUDA events, max over ranks);
`e2e` times the same step through the public API from pinned HOST buffers (H2D of every input + D2H of the result
inside the timed region); `roofline` is the DMoL kernel alone (CUDA events over K launches) against the measured HBM
copy peak; `cpu_baseline` is the reference's own PyTorch implementation of the path (the staged, unmodified reference
under oracle/_ref, written by oracle/make_ref.py) on the host cores, with the C port (oracle/blvm_oracle.c) as a second
figure; `reference_eager_cuda` (N=1) is the same reference code on the same B200 -- the "beat this" number; `sweep`
(N=1) is BASELINE config 5's grid K in {1,10,30} x T in {16000..128000}; N>1 lines carry an `exchange_check` (the fused
NVLink exchange against an NCCL all-reduce, bit for bit) and a `strong` sub-record (a fixed global batch split over the ranks).

Parity anchor of every number here: the reference's code run in fp64 on the same fp32 inputs (tests/parity.py).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "waveform samples/sec, DMoL+KL ELBO fwd+bwd"
UNIT = "samples/s"
NUM_BINS = 65536
BETA, FREE_NATS = 0.5, 0.0625
STRIDE, ZDIM = 64, 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--min-seconds", type=float, default=2.0,
                    help="the K-step timed region is repeated back to back until this much time has passed (so that "
                         "nvidia-smi can observe clocks under load); the median repeat is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reference-device", default="cpu", choices=["cpu", "cpu-torch", "cuda"],
                    help="--impl reference only.  cpu (the contract arm): the C port of the reference path on the host cores.  "
                         "cuda / cpu-torch (informational): the reference's own op chain as eager PyTorch + autograd on this GPU / on "
                         "the host threads (oracle/torch_eager.py, bit-identical to the reference's fp32 run) -- what a blvm "
                         "experiment runs today")
    ap.add_argument("--B", type=int, default=256, help="utterances per GPU")
    ap.add_argument("--T", type=int, default=16000, help="samples per utterance (config 5 sweep: 16000..128000)")
    ap.add_argument("--K", type=int, default=10, help="mixture components (config 5 sweep: 1/10/30)")
    ap.add_argument("--ragged", action="store_true", help="x_sl ~ T*U(0.5,1) instead of full length")
    ap.add_argument("--workload", default="config5", choices=sorted(WORKLOADS),
                    help="config5 (default, the configuration the metric is quoted on; --B/--T/--K apply) or the likelihood/KL/ELBO "
                         "tail of BASELINE configs 2-4 at their full shapes (informational: parity-test cases, not bench lines)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="N=1: skip the config-5 grid (K x T) of DMoL kernel timings")
    ap.add_argument("--no-reference-cuda", action="store_true", help="N=1: skip timing the reference's own eager chain on this GPU")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong-scaling sub-record")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): --B utterances PER GPU; strong: --B utterances in total, split over the ranks (shard_rows)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"],
                    help="element type of the likelihood parameters and their gradient (AMP: the Linear output is consumed as is)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1: scalar-sum exchange fused into the finalize kernel over NVLink peer memory (p2p) or an "
                         "asynchronous NCCL all-reduce (nccl)")
    ap.add_argument("--mode", default="auto", choices=["auto", "graph", "eager"],
                    help="`value` region: replay the step's kernels from CUDA graphs (auto/graph) or call the API eagerly")
    return ap.parse_args()


# latent levels = [(overall stride, width)]; free nats of level l scale with stride_l / stride_0 (clockwork_vae.py:151)
WORKLOADS = {
    "config5": dict(name="config5: standalone DMoL+KL ELBO", levels=[(STRIDE, ZDIM)]),
    "config2": dict(name="config2: WaveNet tail (DMoL only)", B=32, T=16000, K=10, levels=[]),
    "config3": dict(name="config3: SRNN tail", B=64, T=32000, K=10, levels=[(64, 64)]),
    "config3z256": dict(name="config3: SRNN tail, z 256 (benchmarks.txt)", B=64, T=32000, K=10, levels=[(64, 256)]),
    "config4": dict(name="config4: Clockwork-VAE 3-level tail", B=32, T=65536, K=10, levels=[(64, 128), (512, 64), (4096, 32)]),
}


def apply_workload(a):
    w = WORKLOADS[a.workload]
    for k in ("B", "T", "K"):
        if k in w:
            setattr(a, k, w[k])
    a.levels = list(w["levels"])
    return a


def level_free_nats(a, l):
    return FREE_NATS * a.levels[l][0] / a.levels[0][0]


def workload_config(a, n_gpus):
    lv = ", ".join(f"stride {s} width {z}" for s, z in a.levels) or "none"
    kl_bytes = sum(32 * a.B * (a.T // s) * z for s, z in a.levels)
    total = a.B * a.T * 4 * (2 + 6 * a.K) + kl_bytes
    cfg = {
        "workload": f"{WORKLOADS[a.workload]['name']}, per GPU {a.B} utterances x {a.T} samples, DMoL K={a.K}, "
                    f"num_bins={NUM_BINS}, latent layers: {lv}, beta {BETA}, free_nats {FREE_NATS}",
        "utterances_per_gpu": a.B, "samples_per_utterance": a.T, "num_mix": a.K, "num_bins": NUM_BINS,
        "latent_levels": [list(x) for x in a.levels], "ragged": bool(a.ragged),
        "global_samples_per_step": a.B * a.T * n_gpus, "parallelism": f"dp{n_gpus} (utterance sharding)",
        "l2": ("inputs+outputs ~%.2f GB per step per GPU, larger than the 126 MB L2 (no flush needed)" % (total / 1e9)) if total > 2.5e8
              else ("inputs+outputs ~%.0f MB per step per GPU: NOT larger than L2, informational only" % (total / 1e6)),
    }
    if a.workload == "config5":   # keys of the earlier rounds' lines
        cfg.update(latent_stride=STRIDE, latent_width=ZDIM)
    return cfg


def synth_numpy(B, T, K, seed, ragged, levels=((STRIDE, ZDIM),)):
    """Synthetic inputs: y on the rescaled 16-bit grid, raw ~ N(0,1) with locs near y and log-scales
    *2-4 (straddles the -7 clamp and the cdf_delta threshold), KL inputs mu ~ N(0,1), sd = softplus(N(0,1)) + 1e-6."""
    import numpy as np
    rng = np.random.default_rng(seed)
    y = (rng.integers(0, NUM_BINS, (B, T)).astype(np.float32) / np.float32(NUM_BINS - 1) * 2 - 1).astype(np.float32)
    raw = rng.standard_normal((B, T, 3 * K), dtype=np.float32)
    raw[..., K:2 * K] = y[..., None] + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    kls = []
    for stride, zdim in levels:
        kl = [rng.standard_normal((B, T // stride, zdim), dtype=np.float32) for _ in range(4)]
        for i in (1, 3):
            kl[i] = (np.log1p(np.exp(kl[i])) + 1e-6).astype(np.float32)
        kls.append(kl)
    x_sl = np.full(B, T, np.int64)
    if ragged:
        x_sl = (T * rng.uniform(0.5, 1.0, B)).astype(np.int64)
    return y, raw, kls, x_sl


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the C port of the reference path on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_port_throughput(a, steps, warmup, target_step_s=1.5):
    import numpy as np
    from oracle import c_oracle as C
    cores = C.use_all_cores()
    # calibrate on a few utterances, then size the sample so that one step takes ~target_step_s
    rows0 = max(1, min(a.B, cores))
    def lvls(kls):
        return [dict(mu_q=kl[0], sd_q=kl[1], mu_p=kl[2], sd_p=kl[3], stride=a.levels[l][0], free_nats=level_free_nats(a, l))
                for l, kl in enumerate(kls)]

    y, raw, kls, x_sl = synth_numpy(rows0, a.T, a.K, 99, a.ragged, a.levels)
    lv = lvls(kls)
    C.elbo_step(y, raw, x_sl, lv, BETA, a.K, NUM_BINS)
    t0 = time.perf_counter()
    C.elbo_step(y, raw, x_sl, lv, BETA, a.K, NUM_BINS)
    per_row = (time.perf_counter() - t0) / rows0
    rows = int(max(rows0, min(a.B, target_step_s / max(per_row, 1e-9))))
    rows = max(cores, rows - rows % cores) if rows >= cores else rows
    y, raw, kls, x_sl = synth_numpy(rows, a.T, a.K, 100, a.ragged, a.levels)
    lv = lvls(kls)
    for _ in range(warmup):
        C.elbo_step(y, raw, x_sl, lv, BETA, a.K, NUM_BINS)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        C.elbo_step(y, raw, x_sl, lv, BETA, a.K, NUM_BINS)
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    n = float(x_sl.sum())
    sample = (f"{rows} of {a.B} utterances x {a.T} samples (same generator), fp32, {steps} timed steps after {warmup} "
              f"warm-up, C port of the reference path (oracle/blvm_oracle.c: fused loop, closed-form backward), OpenMP")
    return dict(value=n / t, unit=UNIT, cores=cores, kind="port", sample=sample), t


# ----------------------------------------------------------------------------------------------------------------------
# the reference's own implementation of the path (staged under oracle/_ref by oracle/make_ref.py)
# ----------------------------------------------------------------------------------------------------------------------
def reference_step_factory(a, dev, rows, seed=1234):
    """One fwd+bwd of the path made of the REFERENCE'S OWN functions (unmodified, oracle/_ref) on `dev`:
    DiscretizedLogisticMixtureDense.log_prob -> discretized_logistic_mixture_ll (log_likelihoods.py:170-231),
    kl_divergence_gaussian (variational.py:67-70), sequence_mask (operations.py:90-119) and the masked reduction of
    CWVAE.compute_elbo (clockwork_vae.py:132-161: bool masks, free nats scaled per level; any number of levels) or
    WaveNet.compute_loss (wavenet.py:128-146) when there is no latent level, then loss.backward().  Only the 3-line
    split + clamp of distributions.py:383-386 is restated here (in the reference it is fused with the Linear).
    Falls back to the op-for-op restatement oracle/torch_eager.py when the staged reference is absent.
    Returns (step, n_valid_samples, kind, description)."""
    import importlib
    from types import SimpleNamespace

    import torch
    from oracle import ref_loader
    y_np, raw_np, kl_np, x_sl_np = synth_numpy(rows, a.T, a.K, seed, a.ragged, a.levels)
    x_sl = torch.from_numpy(x_sl_np)
    y = torch.from_numpy(y_np).to(dev)
    raw = torch.from_numpy(raw_np).to(dev).requires_grad_(True)
    kls = [[torch.from_numpy(t).to(dev).requires_grad_(True) for t in kl] for kl in kl_np]
    K, T = a.K, a.T

    def zero():
        raw.grad = None
        for kl in kls:
            for t in kl:
                t.grad = None

    if ref_loader.available():
        ref_loader.load()
        dist_mod = importlib.import_module("blvm.modules.distributions")
        var_mod = importlib.import_module("blvm.utils.variational")
        ops_mod = importlib.import_module("blvm.utils.operations")
        cwvae = importlib.import_module("blvm.models.clockwork_vae.clockwork_vae").CWVAE
        wavenet = importlib.import_module("blvm.models.wavenet.wavenet").WaveNet
        lik = dist_mod.DiscretizedLogisticMixtureDense(x_dim=3 * K, y_dim=1, num_mix=K, num_bins=NUM_BINS)
        strides = [s_ for s_, _ in a.levels]
        me = SimpleNamespace(likelihood=lik, num_levels=len(strides), overall_strides=strides)

        def step():
            zero()
            lls = raw[..., K:].view(rows, T, 1, 2 * K)                                   # distributions.py:384
            locs, log_scales = lls.chunk(2, dim=-1)                                     # :385
            params = (raw[..., :K], locs, log_scales.clamp(min=lik.log_epsilon))        # :383, :386
            yy = y.unsqueeze(-1)
            if not strides:
                loss, _, _ = wavenet.compute_loss(me, yy, x_sl, params)
            else:
                klds = [var_mod.kl_divergence_gaussian(*kl) for kl in kls]
                seq_mask = ops_mod.sequence_mask(x_sl, max_len=T, dtype=bool, device=dev)
                level_masks = [ops_mod.sequence_mask((x_sl / s_).ceil().to(int), max_len=T // s_, dtype=bool, device=dev) for s_ in strides]
                loss, _, _, _, _ = cwvae.compute_elbo(me, yy, seq_mask, level_masks, x_sl, params, klds, BETA, FREE_NATS)
            loss.backward()
            return loss

        return step, float(x_sl.sum()), "reference", ("the reference's own functions (oracle/_ref: discretized_logistic_mixture_ll, "
                                                     "kl_divergence_gaussian, sequence_mask, CWVAE.compute_elbo / WaveNet.compute_loss) + autograd")
    from oracle import torch_eager as TE

    def step():
        zero()
        lv = [(*kl, a.levels[l][0], level_free_nats(a, l)) for l, kl in enumerate(kls)]
        loss, _, _, _ = TE.elbo_step(y, raw, x_sl, lv, BETA, K, NUM_BINS, torch.float32)
        loss.backward()
        return loss

    return step, float(x_sl.sum()), "port", "op-for-op eager-PyTorch restatement of the reference chain (oracle/torch_eager.py; oracle/_ref not staged)"


def reference_cpu_throughput(a, steps, warmup, target_step_s=1.5):
    """The reference path as the reference runs it on a CPU: eager PyTorch ops + autograd on all host threads, on a bounded
    sample of the workload's utterances sized so that one step takes about target_step_s."""
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    dev = torch.device("cpu")
    rows0 = max(1, min(a.B, 2))
    step, n0, kind, what = reference_step_factory(a, dev, rows0, seed=99)
    step()
    t0 = time.perf_counter()
    step()
    per_row = (time.perf_counter() - t0) / rows0
    rows = int(max(1, min(a.B, target_step_s / max(per_row, 1e-9))))
    step, n, kind, what = reference_step_factory(a, dev, rows, seed=100)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    sample = (f"{rows} of {a.B} utterances x {a.T} samples (same generator), fp32, {steps} timed steps after {warmup} warm-up, {what}, "
              f"torch.set_num_threads({threads})")
    return dict(value=n / t, unit=UNIT, cores=threads, kind=kind, sample=sample), t


def config1_vrnn_cpu(steps=3, warmup=1):
    """BASELINE config 1 / SURVEY 8d(ii): one full fwd+bwd ELBO step of the reference's VRNNAudio (DMoL-10, 16-bit bins,
    batch 4 x 1 s of 16 kHz audio, frames of 200 samples, h = 256, z = 64) on the host cores: vrnn.py:281-369 end to end."""
    import torch
    from oracle import ref_loader
    if not ref_loader.available():
        return {"unavailable": ref_loader.why_unavailable()}
    ref_loader.load()
    import blvm.models as M
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = M.VRNNAudio(input_size=200, hidden_size=256, latent_size=64, num_mix=10, num_bins=NUM_BINS, likelihood="DMoL")
    model.train()
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, NUM_BINS, (4, 16000), generator=g).float() / (NUM_BINS - 1) * 2 - 1
    x_sl = torch.full((4,), 16000)

    def step():
        model.zero_grad(set_to_none=True)
        loss, _, _ = model(x, x_sl, beta=BETA, free_nats=FREE_NATS)
        loss.backward()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    t = (time.perf_counter() - t0) / steps
    return {"value": 64000 / t, "unit": UNIT, "ms_per_step": t * 1e3, "cores": torch.get_num_threads(), "steps": steps,
            "what": "reference VRNNAudio (DMoL-10, B=4 x 16000 samples, frame 200, h=256, z=64), full model fwd+bwd on the host, unpatched"}


def run_reference_arm(a):
    """Contract arm `--impl reference`: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    budget = 150.0  # seconds for the whole arm
    per_step = min(2.0, budget / (steps + warmup + 2))
    base, t = reference_cpu_throughput(a, steps, warmup, target_step_s=per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a, a.gpus), "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is Python/PyTorch: this arm runs its own, unmodified functions for the path (staged under "
                "oracle/_ref) on all host threads; `cpu_port` is the C/OpenMP port of the same path (a stronger CPU baseline)",
    }
    try:   # second figure: the C port (bounded: a few steps)
        port, _ = cpu_port_throughput(a, steps=min(steps, 5), warmup=min(warmup, 1), target_step_s=min(1.0, per_step))
        line["cpu_port"] = port
    except Exception as ex:
        line["cpu_port"] = {"error": repr(ex)}
    if a.workload == "config5" and not a.no_cpu_baseline:
        try:
            line["config1_vrnn_cpu"] = config1_vrnn_cpu()
        except Exception as ex:
            line["config1_vrnn_cpu"] = {"error": repr(ex)}
    print(json.dumps(line))


def reference_eager_cuda(a, dev, rows=None, steps=10, warmup=3):
    """The reference's own op chain (eager PyTorch kernels + autograd) on this GPU, inputs resident, CUDA events."""
    import torch
    rows = a.B if rows is None else rows
    step, n, kind, what = reference_step_factory(a, dev, rows, seed=1234)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "kind": kind, "what": what,
            "loss": float(loss.detach()), "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9,
            "note": "includes the host sync of the reference's range assert (log_likelihoods.py:195), like every reference step"}


def run_reference_eager_torch(a, on_cuda):
    """Informational arms `--impl reference --reference-device cuda|cpu-torch`."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    if on_cuda:
        rec = reference_eager_cuda(a, torch.device("cuda", 0), steps=max(1, min(a.steps, 50)), warmup=max(3, min(a.warmup, 10)))
        extra = {"peak_memory_gb": rec["peak_memory_gb"], "loss": rec["loss"]}
        value, ms, steps, warmup, where, kind, what = rec["value"], rec["ms_per_step"], rec["steps"], rec["warmup"], "on the same B200", rec["kind"], rec["what"]
    else:
        base, t = reference_cpu_throughput(a, max(1, min(a.steps, 10)), max(1, min(a.warmup, 2)), target_step_s=1.0)
        extra = {"cpu_baseline": base}
        value, ms, steps, warmup, kind, what = base["value"], t * 1e3, max(1, min(a.steps, 10)), max(1, min(a.warmup, 2)), base["kind"], base["sample"]
        where = f"on {base['cores']} host threads"
    print(json.dumps({
        "impl": "reference", "reference_device": "cuda" if on_cuda else "cpu-torch", "metric": METRIC, "value": value,
        "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, 1), "gpu_launches": 0, **extra,
        "reference_kind": kind, "note": f"informational: {what}, {where}",
    }))


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, reasons, mx, power = [], set(), None, []
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    clocks.append(float(f[1]))
                    mx = float(f[2])
                    power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            busy = [c for c, p in zip(clocks, power) if p >= 0.5 * max(power)] or clocks
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(clocks),
                       power_w_max=max(power))
        return out


TORCH_DTYPES = {"f32": "float32", "bf16": "bfloat16", "f16": "float16"}


class DeviceStep:
    """Device-resident inputs of one rank (`rows` utterances) and the step through the public API; optionally replayed from
    two alternating CUDA graphs (separate output workspaces)."""

    def __init__(self, a, dev, rows, seed, exchange=None, pinned=False):
        import torch

        import blvm_b200
        self.a, self.dev, self.rows, self.ex = a, dev, rows, exchange
        y_np, raw_np, kl_np, x_sl_np = synth_numpy(rows, a.T, a.K, seed, a.ragged, a.levels)
        self.x_sl = torch.from_numpy(x_sl_np)
        pin = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
        tdt = getattr(torch, TORCH_DTYPES[a.dtype])
        self.host = dict(y=pin(torch.from_numpy(y_np)), raw=pin(torch.from_numpy(raw_np).to(tdt)),
                         kl=[[pin(torch.from_numpy(t)) for t in kl] for kl in kl_np])
        self.y_d = self.host["y"].to(dev)
        self.raw_d = self.host["raw"].to(dev).requires_grad_(True)
        self.kl_d = [[t.to(dev).requires_grad_(True) for t in kl] for kl in self.host["kl"]]
        self.n_valid = float(self.x_sl.sum())
        self.denom = self.n_valid   # per-rank normaliser; the global loss is recombined from the exchanged sums
        self.params = blvm_b200.DMoLParams(self.raw_d, a.K, 1, -7.0)
        self.x_dev = self.x_sl.to(dev)       # `value` arm: every input, the lengths included, is resident in HBM
        self.lens_dev = [blvm_b200.level_lengths(self.x_dev, s) for s, _ in a.levels]
        self.fnats = [level_free_nats(a, l) for l in range(len(a.levels))]
        self.scaler = None
        if a.dtype == "f16":   # fp16 gradients need the GradScaler's factor inside the kernel (amp.py): the reference's AMP loop
            self.scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
            self.scaler._lazy_init_scale_growth_tracker(dev)
        self.graphs, self.i, self.launches_per_step = None, 0, None

    def compute(self):
        """One step of the path through the public API: forward (values + gradients) and backward."""
        import blvm_b200
        self.raw_d.grad = None
        for kl in self.kl_d:
            for t in kl:
                t.grad = None
        levels = [blvm_b200.KLLevel(*kl, lens=self.lens_dev[l], free_nats=self.fnats[l]) for l, kl in enumerate(self.kl_d)]
        out = blvm_b200.fused_elbo(self.y_d, self.params, self.x_sl, levels, BETA, FREE_NATS, num_bins=NUM_BINS, denom=self.denom,
                                   x_sl_device=self.x_dev, exchange=self.ex, grad_scaler=self.scaler)
        if self.scaler is not None:
            self.scaler.scale(out.loss).backward()
        else:
            out.loss.backward()
        # with `exchange=ex` the finalize kernel itself publishes this step's sums to every rank over NVLink peer memory
        # and adds up the previous step's slots into ex.global_sums: the exchange costs no launch and no host call
        return out.sums

    def capture(self):
        """Two CUDA graphs of the step (likelihood, KL of all levels, finalize), replayed alternately.  The kernels launched
        while capturing are COUNTED (ops.launch_count): that is the number of launches every replay performs."""
        import torch
        from blvm_b200 import ops
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self.compute()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for _ in range(2):
            g = torch.cuda.CUDAGraph()
            ops.reset_launch_count()
            with torch.cuda.graph(g):
                sums_static = self.compute()
            self.launches_per_step = ops.launch_count()
            graphs.append((g, sums_static))
        torch.cuda.synchronize()
        self.graphs = graphs

    def replay(self):
        g, sums_static = self.graphs[self.i & 1]
        self.i += 1
        g.replay()
        return sums_static


def timed_region(a, runner, steps, sync_all, world, dev, min_seconds):
    """EXACTLY `steps` steps between two CUDA events, barrier + synchronize on both sides; repeated back to back until
    min_seconds have passed (so that nvidia-smi can see clocks under load); returns the list of region times (ms)."""
    import torch
    import torch.distributed as dist
    region_ms = []
    t_begin = time.perf_counter()
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(steps):
            runner()
        e1.record()
        sync_all()
        region_ms.append(e0.elapsed_time(e1))
        done = torch.tensor([1.0 if (time.perf_counter() - t_begin >= min_seconds or len(region_ms) >= 200) else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(done, op=dist.ReduceOp.MAX)   # all ranks leave the loop together
        if done.item() > 0:
            return region_ms


def dmol_kernel_us(dev, B, T, K, dtype, iters):
    """The dominant kernel alone (value + gradient + masked row partials): `iters` launches between two CUDA events on the
    launching stream.  Returns (microseconds per launch, algorithmic bytes per launch)."""
    import torch

    import blvm_b200
    from blvm_b200 import ops
    tdt = getattr(torch, TORCH_DTYPES[dtype])
    esz = 4 if dtype == "f32" else 2
    g = torch.Generator(device=dev).manual_seed(7)
    y = (torch.randint(0, NUM_BINS, (B, T), device=dev, generator=g).float() / (NUM_BINS - 1) * 2 - 1)
    raw = torch.randn(B, T, 3 * K, device=dev, generator=g)
    raw[..., K:2 * K] = y.unsqueeze(-1) + 0.1 * raw[..., K:2 * K]
    raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
    raw = raw.to(tdt)
    x_dev = torch.full((B,), T, dtype=torch.int64, device=dev)
    part = torch.empty(B * int(blvm_b200._lib.lib.blvm_dmol_chunks(T, K, 1)), dtype=torch.float64, device=dev)
    graw = torch.empty_like(raw)
    gs = -1.0 / float(B * T) * (65536.0 if dtype == "f16" else 1.0)

    def launch():
        ops._dmol_call(y, raw, x_dev, None, gs, B, T, K, 1, NUM_BINS, -7.0, 1, None, graw, part)

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(iters):
        launch()
    k1.record()
    torch.cuda.synchronize()
    us = k0.elapsed_time(k1) / iters * 1e3
    del raw, graw, part, y
    return us, B * T * (8 + 6 * K * esz)


def fused_head_record(dev, B, T, K, peak, iters=20):
    """the likelihood head -- nn.Linear(30 -> 3K) + DMoL value / gradient + the Linear's backward -- as ONE
    tcgen05 tensor-core kernel (`DiscretizedLogisticMixtureDense(fuse_linear=True)` -> fused_elbo -> backward) against the same head
    unfused (cuBLAS Linear, the DMoL kernel on its bf16 output, autograd's two backward GEMMs), both through the public API."""
    import torch

    import blvm_b200
    Din = 3 * K
    g = torch.Generator(device=dev).manual_seed(11)
    y = (torch.randint(0, NUM_BINS, (B, T), device=dev, generator=g).float() / (NUM_BINS - 1) * 2 - 1)
    x = torch.randn(B, T, Din, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    x_sl = torch.full((B,), T, dtype=torch.int64)
    x_dev = x_sl.to(dev)
    out = {}
    for name, fuse in (("fused", True), ("unfused", False)):
        lik = blvm_b200.DiscretizedLogisticMixtureDense(Din, 1, K, NUM_BINS, fuse_linear=fuse).to(dev)
        with torch.no_grad():
            lik.params.bias[2 * K:] -= 4.0

        def step():
            x.grad = None
            lik.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                params = lik(x)
            r = blvm_b200.fused_elbo(y, params, x_sl, (), num_bins=NUM_BINS, denom=float(B * T), x_sl_device=x_dev)
            r.loss.backward()
            return r

        # replayed from a CUDA graph so that the GPU time is measured, not the Python around it (eager: see `eager_us`)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            r = step()
        e1.record()
        torch.cuda.synchronize()
        eager_us = e0.elapsed_time(e1) / iters * 1e3
        # the last eager result keeps its autograd graph (and the leaves' gradient accumulators, bound to the stream they were created
        # on: here the legacy default stream) alive; a capture that reuses them fails with cudaErrorStreamCaptureImplicit
        del r
        x.grad = None
        lik.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            r = step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"us_per_step": e0.elapsed_time(e1) / iters * 1e3, "eager_us": eager_us, "loss": float(r.loss)}
    byts = B * T * (8 + 4 * Din)       # x and dx in bf16, y and log-prob in fp32: 128 B/sample at x_dim 30
    us = out["fused"]["us_per_step"]
    return {"what": f"likelihood head fwd+bwd, bf16 activations, x_dim {Din} -> 3K = {3 * K}, B={B} x T={T}: Linear + DMoL + Linear backward in one "
                    "tcgen05 kernel (+ dW reduce + finalize) vs cuBLAS Linear + DMoL kernel + autograd GEMMs, public API, CUDA events",
            "fused_us": us, "unfused_us": out["unfused"]["us_per_step"], "speedup": out["unfused"]["us_per_step"] / us,
            "fused_eager_us": out["fused"]["eager_us"], "unfused_eager_us": out["unfused"]["eager_us"],
            "note": "the two losses differ in the 3rd digit because the unfused AMP path rounds the Linear output to bf16 before the likelihood "
                    "reads it; the fused kernel keeps the fp32 accumulators (tests/test_gpu_linear_head.py compares both with an fp32 evaluation)",
            "algorithmic_bytes": byts, "achieved_gbs": byts / us / 1e3, "frac_of_hbm_peak": byts / us / 1e3 / peak,
            "loss_fused": out["fused"]["loss"], "loss_unfused": out["unfused"]["loss"]}


def run_gpu_arm(a):
    import torch
    import torch.distributed as dist

    import blvm_b200
    from blvm_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a GPU (blvm_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # before any pinned allocation: keep this rank (and, by first touch, its host buffers) on the GPU's NUMA node
    numa_node = blvm_b200.bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    T, K = a.T, a.K
    strong_main = a.scaling == "strong" and world > 1
    if strong_main:
        lo, hi = blvm_b200.shard_rows(a.B, rank, world)
        B = hi - lo
    else:
        B = a.B

    ex = None
    if world > 1 and a.exchange in ("auto", "p2p"):
        try:
            ex = blvm_b200.SumsExchange()
        except Exception as err:   # no symmetric memory on this stack: NCCL all-reduce instead
            if a.exchange == "p2p":
                raise
            sys.stderr.write(f"[bench] symmetric-memory exchange unavailable ({err!r}); using NCCL\n")
            ex = None
    exchange_kind = "none" if world == 1 else ("p2p (fused into finalize, NVLink peer stores)" if ex is not None else "nccl all-reduce (async)")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    pending = []

    def make_runner(stepper, mode_req):
        """Returns (runner, mode).  N>1 without the fused exchange: an asynchronous NCCL all-reduce of the 8 sums per step."""
        def exchange(sums):
            if world > 1 and ex is None:
                pending.append(blvm_b200.all_reduce_sums(sums, async_op=True, inplace=True))

        def step_eager():
            while pending:
                pending.pop(0).wait()
            exchange(stepper.compute())

        if mode_req in ("auto", "graph"):
            try:
                stepper.capture()

                def step_graph():
                    while len(pending) >= 2:       # this graph's previous exchange must be done before it rewrites its sums
                        pending.pop(0).wait()
                    exchange(stepper.replay())
                return step_graph, "graph"
            except Exception as err:  # capture not possible on this stack: fall back to the eager step
                if mode_req == "graph":
                    raise
                sys.stderr.write(f"[bench] CUDA graph capture failed ({err!r}); running eagerly\n")
        return step_eager, "eager"

    def drain():
        while pending:
            pending.pop().wait()

    main = DeviceStep(a, dev, B, 1234 + rank, exchange=ex, pinned=not a.no_e2e)
    runner, mode = make_runner(main, a.mode)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # nvidia-smi needs ~1 s before its first sample: start it ahead of the warm-up
    for _ in range(max(a.warmup, 3)):
        runner()
    drain()
    sync_all()
    ops.reset_launch_count()

    def timed_runner():
        runner()
    region_ms = timed_region(a, timed_runner, a.steps, lambda: (drain(), sync_all()), world, dev, a.min_seconds)
    ms_total = statistics.median(region_ms)
    if mode == "eager":
        launches = ops.launch_count() // max(len(region_ms), 1)
    else:
        launches = (main.launches_per_step or 0) * a.steps    # counted while capturing; every replay launches exactly these

    # ---- N > 1: the fused exchange must deliver the right numbers ----------------------------------------------
    exchange_check = None
    if world > 1:
        sums_local = main.compute().clone()                   # one more step, eagerly: its sums ...
        torch.cuda.synchronize()
        ref_global = blvm_b200.combine_sums(blvm_b200.all_reduce_sums(sums_local), BETA)   # ... all-reduced by NCCL (the checker)
        if ex is not None:
            got = ex.consume(beta=BETA, lag=0)                # ... and as the NVLink peer-store all-gather delivered them
            torch.cuda.synchronize()
            ex.check()
            # both add the same W fp64 values per entry, ours in rank order, NCCL in its ring / tree order: equal up to the last bits
            # (bit-identical at W = 2 and 4 on this pool, 1e-14 relative at W = 8)
            diff = (got[:7] - ref_global[:7]).abs()
            rel = (diff / ref_global[:7].abs().clamp_min(1e-300)).max()
            flag = torch.tensor([float(diff.max()), float(rel)], device=dev, dtype=torch.float64)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            exchange_check = {"ok": bool(flag[1].item() <= 1e-9), "max_abs_diff": float(flag[0].item()), "max_rel_diff": float(flag[1].item()),
                              "step": int(got[7].item()), "global_loss": float(got[0].item()),
                              "what": "exchange.consume(lag=0) vs NCCL all-reduce of the same step's sums (fp64), entries 0-6, max over ranks; ok = rel <= 1e-9 (observed: 0 at 2 and 4 ranks, 1e-14 at 8)"}
            if not exchange_check["ok"]:
                raise SystemExit(f"[bench] fused exchange delivered wrong sums: {exchange_check}")
        else:
            exchange_check = {"ok": True, "max_abs_diff": 0.0, "what": "NCCL all-reduce is the exchange on this run"}

    # ---- the dominant kernel alone ------------------------------------------------------------------------------------------
    kern_us, alg_bytes = dmol_kernel_us(dev, B, T, K, a.dtype, max(a.steps, 200))
    kern_ms = kern_us * 1e-3
    # context for the roofline: the same device's copy bandwidth measured the same way (back-to-back launches for ~1 s, i.e.
    # under the same power cap as the kernel loop above), next to the burst figure of MEASURED_PEAKS.json
    sustained_copy = None
    if rank == 0:
        try:
            src = torch.empty(1 << 28, dtype=torch.float32, device=dev)     # 1 GiB, read + written: 2 GiB per copy
            dst = torch.empty_like(src)
            for _ in range(3):
                dst.copy_(src)
            torch.cuda.synchronize()
            n_copy = 600
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(n_copy):
                dst.copy_(src)
            c1.record()
            torch.cuda.synchronize()
            sustained_copy = 2 * src.numel() * 4 * n_copy / (c0.elapsed_time(c1) * 1e-3) / 1e9
            del src, dst
        except Exception:
            sustained_copy = None
    clocks = sampler.stop() if rank == 0 else None

    # ---- strong scaling (N > 1): the SAME global batch of a.B utterances split over the ranks ------------------------------
    strong = None
    if world > 1 and not a.no_strong and not strong_main:
        lo, hi = blvm_b200.shard_rows(a.B, rank, world)
        st = DeviceStep(a, dev, hi - lo, 4321 + rank, exchange=ex)
        s_runner, s_mode = make_runner(st, a.mode)
        for _ in range(max(a.warmup, 3)):
            s_runner()
        drain()
        s_ms = statistics.median(timed_region(a, s_runner, a.steps, lambda: (drain(), sync_all()), world, dev, min(a.min_seconds, 1.0)))
        tt = torch.tensor([s_ms, st.n_valid], device=dev, dtype=torch.float64)
        mx, sm = tt.clone(), tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        s_step = float(mx[0]) / a.steps
        strong = {"scaling": "strong", "value": float(sm[1]) / (s_step * 1e-3), "unit": UNIT, "ms_per_step": s_step, "mode": s_mode,
                  "global_utterances": a.B, "utterances_per_gpu": hi - lo,
                  "what": f"the N=1 workload ({a.B} utterances x {T}) split over {world} ranks by shard_rows; no data-path collective, "
                          f"exchange: {exchange_kind}; compare with the N=1 line's value for the strong-scaling speed-up"}
        del st

    # ---- e2e: the public API from pinned host buffers, H2D + D2H inside the timed region -----------------------------
    e2e = None
    if not a.no_e2e:
        host, x_sl = main.host, main.x_sl
        in_tensors = [host["y"], host["raw"], *[t for kl in host["kl"] for t in kl]]
        h2d = sum(t.numel() * t.element_size() for t in in_tensors) + x_sl.numel() * 8
        d2h = 8 * 8 + 4 * B * 8

        def step_e2e():
            y = host["y"].to(dev, non_blocking=True)
            raw = host["raw"].to(dev, non_blocking=True).requires_grad_(True)
            kls = [[t.to(dev, non_blocking=True).requires_grad_(True) for t in kl] for kl in host["kl"]]
            levels = [blvm_b200.KLLevel(*kl, stride=a.levels[l][0], free_nats=main.fnats[l]) for l, kl in enumerate(kls)]
            out = blvm_b200.fused_elbo(y, blvm_b200.DMoLParams(raw, K, 1, -7.0), x_sl, levels,
                                       BETA, FREE_NATS, num_bins=NUM_BINS, denom=main.denom, grad_scaler=main.scaler)
            (main.scaler.scale(out.loss) if main.scaler is not None else out.loss).backward()
            sums = out.sums
            if world > 1:
                sums = blvm_b200.all_reduce_sums(sums)
            return sums.cpu(), torch.stack([out.log_prob, out.kl, out.kl_fn, out.elbo]).cpu()  # D2H (synchronises)

        n_e2e = max(3, min(a.steps, 10))
        for _ in range(2):
            step_e2e()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_e2e()
        sync_all()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        # the ceiling of this arm: the bare pinned-memory copies of the same buffers, nothing else, all ranks at once
        dsts = [torch.empty_like(t, device=dev) for t in in_tensors]
        for _ in range(2):
            for d_, s_ in zip(dsts, in_tensors):
                d_.copy_(s_, non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            for d_, s_ in zip(dsts, in_tensors):
                d_.copy_(s_, non_blocking=True)
            torch.cuda.synchronize()
        sync_all()
        copy_s = (time.perf_counter() - t0) / n_e2e
        del dsts
        tt = torch.tensor([e2e_s, copy_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s, copy_s = float(tt[0]), float(tt[1])
        e2e = {"value": main.n_valid * n_gpus / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s * 1e3, "steps": n_e2e, "h2d_gbs_per_gpu": h2d / e2e_s / 1e9,
               "pinned_copy_ceiling": {"h2d_gbs_per_gpu": h2d / copy_s / 1e9, "ms_per_step": copy_s * 1e3,
                                       "value": main.n_valid * n_gpus / copy_s,
                                       "what": "the same pinned buffers copied host->device and nothing else, all ranks concurrently, max over ranks"},
               "frac_of_copy_ceiling": copy_s / e2e_s}

    # ---- max over ranks, totals --------------------------------------------------------------------------------------
    tot = torch.tensor([ms_total, kern_ms, main.n_valid], device=dev, dtype=torch.float64)
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, kern_ms, n_global = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        n_global = main.n_valid
    ms_step = ms_total / a.steps

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):   # DRAM bytes of ONE `ncu --set full` capture of this build's kernel (not of this run: ncu cannot run inside a timed bench)
            try:
                rec = json.load(open(tp)).get(f"dmol_fwd_grad_K{K}_B{B}_T{T}_{a.dtype}", {})
                traffic, traffic_src = rec.get("dram_bytes_per_launch"), rec.get("source")
            except Exception:
                traffic = None
        cfg = workload_config(a, n_gpus)
        cfg["utterances_per_gpu"] = B
        cfg["global_samples_per_step"] = int(n_global) if not a.ragged else cfg["global_samples_per_step"]
        line = {
            "metric": METRIC, "value": n_global / (ms_step * 1e-3), "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong_main else "weak",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "kernel": f"dmol_tile_kernel<K={K},128,grad,{a.dtype}>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": kern_ms * 1e3,
                         "note": "algorithmic bytes = SURVEY 8d's 4(2+6K) B/sample (fp32; 16-bit parameters: 8+12K); the fused step reduces the "
                                 "per-sample log-prob to row partials instead of writing it, so the kernel is REQUIRED to move 4 B/sample less "
                                 "(1.6 % at K=10) -- `traffic` is what it actually moved",
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "sustained_copy_gbs": sustained_copy,
                         "frac_of_sustained_copy": (achieved / sustained_copy) if sustained_copy else None},
            "e2e": e2e, "numa_node_rank0": numa_node, "gpu_launches": launches, "clocks": clocks, "mode": mode, "exchange": exchange_kind,
            "timed_region_repeats": len(region_ms), "timed_region_ms_min_max": [min(region_ms), max(region_ms)],
            "parity_anchor": "the reference's code run in fp64 on the same fp32 inputs (tests/parity.py): values/gradients 1e-5, sums 1e-6; "
                             "the reference's own fp32 run is up to 1e-3 away from that on 16-bit audio (sigmoid(a) - sigmoid(b) cancellation)",
            "step": "fused_elbo(...).loss.backward() through the Python API" + (f" ({main.launches_per_step} kernels per step, counted during capture: likelihood, KL of all levels, finalize; replayed from CUDA graphs)" if mode == "graph" else "")
                    + (("; sums exchange: " + exchange_kind) if world > 1 else ""),
        }
        if exchange_check is not None:
            line["exchange_check"] = exchange_check
        if strong is not None:
            line["strong"] = strong
        if n_gpus == 1 and not a.no_sweep and a.workload == "config5":
            # BASELINE config 5 is a sweep: the dominant kernel at every grid point, measured in this run
            sweep = []
            for Ks in (1, 10, 30):
                for Ts in (16000, 32000, 64000, 128000):
                    try:
                        us, byts = dmol_kernel_us(dev, a.B, Ts, Ks, a.dtype, 30)
                        sweep.append({"K": Ks, "T": Ts, "us": us, "GBs": byts / us / 1e3, "frac": byts / us / 1e3 / peak})
                    except Exception as err:
                        sweep.append({"K": Ks, "T": Ts, "error": repr(err)})
                    torch.cuda.empty_cache()
            line["sweep"] = {"what": f"dmol fwd+grad kernel alone, B={a.B}, {a.dtype} parameters, 30 launches between CUDA events, frac = algorithmic GB/s / peak; "
                                     "as in the fused step the per-sample log-prob is reduced to row partials and not written (the algorithmic figure of "
                                     "SURVEY 8d, 4(2+6K) B/sample, counts those 4 bytes: 1.6 % at K=10, 12.5 % at K=1)",
                             "points": sweep}
            # the other instantiated mixture sizes at T = 16000 (stream kernel K <= 5, rotated section walk at K = 8 / 16)
            other = []
            for Ks in (2, 3, 4, 5, 8, 12, 16, 20):
                try:
                    us, byts = dmol_kernel_us(dev, a.B, 16000, Ks, a.dtype, 30)
                    other.append({"K": Ks, "T": 16000, "us": us, "GBs": byts / us / 1e3, "frac": byts / us / 1e3 / peak})
                except Exception as err:
                    other.append({"K": Ks, "T": 16000, "error": repr(err)})
                torch.cuda.empty_cache()
            line["sweep"]["other_K"] = other
        if n_gpus == 1 and not a.no_sweep and a.workload == "config5" and K == 10:
            try:
                line["fused_head"] = fused_head_record(dev, a.B, T, K, peak)
            except Exception as err:
                line["fused_head"] = {"error": repr(err)}
            torch.cuda.empty_cache()
        if n_gpus == 1 and not a.no_reference_cuda:
            try:   # the "beat this" number: the reference's own eager chain on this same GPU, same shape
                line["reference_eager_cuda"] = reference_eager_cuda(a, dev)
                line["reference_eager_cuda"]["ratio_device_timed"] = line["value"] / line["reference_eager_cuda"]["value"]
            except Exception as err:
                line["reference_eager_cuda"] = {"error": repr(err)}
            torch.cuda.empty_cache()
        if not a.no_cpu_baseline and n_gpus == 1:
            try:
                line["cpu_baseline"], _ = reference_cpu_throughput(a, steps=8, warmup=1, target_step_s=1.0)   # ~10-15 s of CPU work
            except Exception as err:  # the checker must never take the measurement down
                line["cpu_baseline"] = {"error": repr(err)}
            try:
                line["cpu_port"], _ = cpu_port_throughput(a, steps=10, warmup=1, target_step_s=0.4)
            except Exception as err:
                line["cpu_port"] = {"error": repr(err)}
            if a.workload == "config5":
                try:
                    line["config1_vrnn_cpu"] = config1_vrnn_cpu()
                except Exception as err:
                    line["config1_vrnn_cpu"] = {"error": repr(err)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = apply_workload(parse())
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 on their own (NCCL prints its version banner
    # there) are pointed at stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if a.impl == "reference" and a.reference_device != "cpu":
        run_reference_eager_torch(a, a.reference_device == "cuda")
    elif a.impl == "reference":
        run_reference_arm(a)
    else:
        run_gpu_arm(a)


if __name__ == "__main__":
    main()

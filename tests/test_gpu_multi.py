"""Multi-GPU tests (need >= 2 GPUs on one node: run under `gpurun --gpus 2`; skipped on the single-GPU box).
Two ranks (one process per GPU, NCCL for rendezvous) run the fused ELBO on their shard with the scalar exchange fused
into the finalize kernel over NVLink peer memory; the consumed global sums must equal the single-process values."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import blvm_b200 as B
    g = load_golden("elbo_srnn_a")
    K, nb, S = int(g["K"]), int(g["num_bins"]), int(g["stride"])
    beta, fn = float(g["beta"]), float(g["free_nats"])
    lo, hi = B.shard_rows(len(g["x_sl"]), rank, world)
    dev = torch.device("cuda", rank)
    cu = lambda a: torch.as_tensor(np.asarray(a)[lo:hi]).float().to(dev)
    x_sl = torch.as_tensor(g["x_sl"][lo:hi])
    denom = B.global_denominator(x_sl)
    ex = B.SumsExchange()
    results = []
    for step in range(6):      # several steps: exercises the 4-deep slot ring and the lagged consume
        raw = cu(g["raw"]).requires_grad_(True)
        ins = [cu(g[n]).requires_grad_(True) for n in ("mu_q", "sd_q", "mu_p", "sd_p")]
        out = B.fused_elbo(cu(g["y"]), B.DMoLParams(raw, K, 1, -7.0), x_sl, [B.KLLevel(*ins, stride=S)], beta, fn,
                           num_bins=nb, denom=denom, exchange=ex)
        out.loss.backward()
        results.append(ex.global_sums.clone())     # written by the finalize kernel: the previous step's global sums
    last = ex.consume(beta=beta, lag=0)
    torch.cuda.synchronize()
    ex.check()
    # NCCL path for comparison
    ref = B.combine_sums(B.all_reduce_sums(out.sums), beta)
    q.put((rank, [r.cpu().numpy() for r in results], last.cpu().numpy(), ref.cpu().numpy(), out.loss.item(), denom))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
@pytest.mark.timeout(300)
def test_fused_finalize_exchange_two_gpus():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = load_golden("elbo_srnn_a")
    for rank, results, last, ref, loss_local, denom in got:
        assert results[0][7] == 0                      # lag 1: nothing to consume after the first step
        for i, r in enumerate(results[1:], start=1):
            assert r[7] == i                           # step i consumed after step i+1 was published
        assert last[7] == 6
        for r in results[1:] + [last]:
            np.testing.assert_allclose(r[0], g["loss64"], rtol=1e-6)          # global loss on every rank
            np.testing.assert_allclose(r[4], g["elbo64"].sum(), rtol=1e-6)
            np.testing.assert_allclose(r[5], g["x_sl"].sum(), rtol=0)
            np.testing.assert_allclose(r[6], -g["elbo64"].sum() / np.log(2) / g["x_sl"].sum(), rtol=1e-6)
        np.testing.assert_allclose(last[:7], ref[:7], rtol=1e-12)             # identical to the NCCL all-reduce path
    assert np.array_equal(got[0][2], got[1][2])                              # bit-identical on both ranks
    np.testing.assert_allclose(np.mean([x[4] for x in got]), g["loss64"], rtol=1e-6)   # mean of rank losses == global


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_one_process_drives_two_devices():
    """Per-device launch state: function attributes (opt-in dynamic shared memory of the stream / tile kernels),
    occupancy and SM count are cached per CUDA device.  One process runs the same small-K (stream kernel: > 48 KB of dynamic
    shared memory at K = 5) and K = 30 (tile kernel: 46 KB slabs) steps on cuda:0 first and then on cuda:1; results must be
    bit-identical and the second device must not fail with an invalid-argument launch error."""
    import blvm_b200 as B
    g = torch.Generator().manual_seed(11)
    for K in (5, 4, 30, 10):
        Bn, T, nb = 3, 4096, 65536
        y = (torch.randint(0, nb, (Bn, T), generator=g).float() / (nb - 1) * 2 - 1)
        raw = torch.randn(Bn, T, 3 * K, generator=g)
        raw[..., 2 * K:] = raw[..., 2 * K:] * 2 - 4
        x_sl = torch.tensor([T, T - 100, T // 2])
        outs = []
        for d in (0, 1):
            dev = torch.device("cuda", d)
            r = raw.to(dev).requires_grad_(True)
            out = B.fused_elbo(y.to(dev), B.DMoLParams(r, K, 1, -7.0), x_sl, (), num_bins=nb)
            out.loss.backward()
            s, m, _ = B.ops.dmol_sample_mode(r, K, 1, -7.0)
            torch.cuda.synchronize(dev)
            outs.append((out.loss.item(), out.log_prob.cpu(), r.grad.cpu(), m.cpu()))
        assert outs[0][0] == outs[1][0]
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][3], outs[1][3])


def _ddp_worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import blvm_b200 as B
    model, x, y, x_sl, S = _ddp_problem(dev)
    ddp = DDP(model, device_ids=[rank])
    lo, hi = B.shard_rows(x.shape[0], rank, world)
    denom = B.global_denominator(x_sl[lo:hi])           # sum over ALL ranks / world: the mean of the rank losses is the global loss
    loss = _ddp_loss(B, ddp, x[lo:hi], y[lo:hi], x_sl[lo:hi], S, denom)
    loss.backward()                                      # DDP averages the gradients over the ranks (NCCL over NVLink)
    grads = {n: p.grad.detach().double().cpu() for n, p in model.named_parameters()}
    losses = [torch.zeros((), dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.detach().double())
    q.put((rank, grads, float(torch.stack(losses).mean())))
    dist.barrier()
    dist.destroy_process_group()


def _ddp_problem(dev):
    """A latent-variable head in miniature: h -> (DMoL parameters, Gaussian posterior / prior parameters); seeded identically everywhere."""
    import blvm_b200 as B
    torch.manual_seed(21)
    Bn, T, H, S, Z = 6, 1024, 24, 64, 8

    class Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Linear(H, H)
            self.likelihood = B.DiscretizedLogisticMixtureDense(H, 1, 10, 65536)
            self.latent = torch.nn.Linear(H, 4 * Z)

        def forward(self, x):
            h = torch.tanh(self.body(x))
            q = self.latent(h[:, ::S])
            mu_q, sd_q, mu_p, sd_p = q.chunk(4, -1)
            return self.likelihood(h), (mu_q, torch.nn.functional.softplus(sd_q) + 1e-3, mu_p, torch.nn.functional.softplus(sd_p) + 1e-3)

    model = Head().to(dev)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(Bn, T, H, generator=g).to(dev)
    y = (torch.randint(0, 65536, (Bn, T), generator=g).float() / 65535 * 2 - 1).to(dev)
    x_sl = torch.tensor([1024, 900, 777, 512, 300, 64])      # very different lengths per shard: the denominators matter
    return model, x, y, x_sl, S


def _ddp_loss(B, model, x, y, x_sl, S, denom):
    params, kl = model(x)
    return B.fused_elbo(y, params, x_sl, [B.KLLevel(*kl, stride=S)], 0.5, 0.25, num_bins=65536, denom=denom).loss


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
@pytest.mark.timeout(300)
def test_ddp_gradients_equal_the_single_process_step():
    """utterances sharded over two ranks, model wrapped in DistributedDataParallel (NCCL over NVLink).  With
    `denom = global_denominator(x_sl_local)` (= sum_global(x_sl) / world) the MEAN of the rank losses is the single-process loss of
    vrnn.py:277 and DDP's averaged gradients equal the single-process gradients, although the shards hold very different numbers of
    valid samples."""
    import torch.multiprocessing as mp
    import blvm_b200 as B
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    model, x, y, x_sl, S = _ddp_problem(dev)
    loss = _ddp_loss(B, model, x, y, x_sl, S, None)          # the whole batch in one process: denom = sum(x_sl)
    loss.backward()
    np.testing.assert_allclose(got[0][2], float(loss), rtol=1e-6)            # mean of the rank losses == global loss
    for n, p in model.named_parameters():
        ref = p.grad.detach().double().cpu()
        scale = float(ref.abs().max()) + 1e-30
        for rank, grads, _ in got:
            assert float((grads[n] - ref).abs().max()) <= 2e-5 * scale, (n, rank)     # fp32 matmul reduction order differs between B=3 and B=6
        assert torch.equal(got[0][1][n], got[1][1][n])                        # identical on both ranks after the all-reduce

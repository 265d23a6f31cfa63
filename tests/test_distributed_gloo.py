"""The N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding and the scalar-sum exchange
(blvm_b200.distributed).  The per-rank 'kernel' here is the numpy oracle — this test is about the exchange, i.e. that
sharded sums + all-reduce + combine reproduce the single-process ELBO numbers, and that denom = global/world makes the
mean of the rank losses the global loss."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from blvm_b200.distributed import all_reduce_sums, combine_sums, global_denominator, shard_rows
    from oracle import blvm_oracle as O

    g = load_golden("elbo_srnn_a")
    B = len(g["x_sl"])
    lo, hi = shard_rows(B, rank, world)
    beta, fn, S = float(g["beta"]), float(g["free_nats"]), int(g["stride"])
    x_sl = g["x_sl"][lo:hi]
    denom = global_denominator(torch.as_tensor(x_sl))          # global sum / world
    lv = [dict(mu_q=g["mu_q"][lo:hi], sd_q=g["sd_q"][lo:hi], mu_p=g["mu_p"][lo:hi], sd_p=g["sd_p"][lo:hi], stride=S,
               free_nats=fn)]
    r = O.fused_elbo_value_and_grad(g["y"][lo:hi], g["raw"][lo:hi], x_sl, lv, beta, int(g["K"]), int(g["num_bins"]))
    # local loss normalised by denom (what fused_elbo(denom=...) returns) and its gradient
    local_sum = float(x_sl.sum())
    loss_local = r["loss"] * local_sum / denom
    graw_local = r["graw"] * local_sum / denom
    sums = torch.tensor([loss_local, r["logp"].sum(), r["kl"].sum(), r["kl_fn"].sum(), r["elbo"].sum(), local_sum, 0.0, 0.0],
                        dtype=torch.float64)
    pend = all_reduce_sums(sums, async_op=True)
    tot = combine_sums(pend.wait(), beta)
    # DDP averages gradients: emulate with an all-reduce(mean) of the (zero-padded) gradient
    full = np.zeros_like(g["graw64"])
    full[lo:hi] = graw_local
    t = torch.from_numpy(full)
    dist.all_reduce(t)
    t /= world
    mean_loss = torch.tensor([loss_local], dtype=torch.float64)
    dist.all_reduce(mean_loss)
    mean_loss /= world
    if rank == 0:
        q.put(dict(tot=tot.numpy(), grad=t.numpy(), mean_loss=float(mean_loss), denom=denom, shard=(lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_sums_and_ddp_normalisation():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = load_golden("elbo_srnn_a")
    assert res["denom"] == g["x_sl"].sum() / world
    np.testing.assert_allclose(res["tot"][0], g["loss64"], rtol=1e-12)            # global loss from reduced sums
    np.testing.assert_allclose(res["tot"][4], g["elbo64"].sum(), rtol=1e-12)
    np.testing.assert_allclose(res["tot"][5], g["x_sl"].sum(), rtol=0)
    np.testing.assert_allclose(res["tot"][6], -g["elbo64"].sum() / np.log(2) / g["x_sl"].sum(), rtol=1e-12)
    np.testing.assert_allclose(res["mean_loss"], g["loss64"], rtol=1e-12)         # mean of rank losses == global loss
    assert np.abs(res["grad"] - g["graw64"]).max() / np.abs(g["graw64"]).max() < 1e-9  # DDP-averaged grads == global grads


def test_shard_rows_partition():
    from blvm_b200.distributed import shard_rows
    for n in (0, 1, 7, 256, 257):
        for w in (1, 2, 3, 8):
            parts = [shard_rows(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1

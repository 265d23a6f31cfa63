"""Batch-sharded data parallelism for the path: utterances are independent, so each rank runs the
kernels on its own B/W rows; the only exchange is one all-reduce of the scalar sums (NCCL over NVLink on GPUs, gloo in
the CPU tests).  No collective touches per-sample data."""
from typing import Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_rows", "global_denominator", "all_reduce_sums", "combine_sums"]


def shard_rows(n_rows: int, rank: int, world_size: int):
    """Contiguous [start, stop) rows of rank `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_denominator(x_sl_local: torch.Tensor, group=None) -> float:
    """sum(x_sl) over all ranks divided by world_size: pass it as `fused_elbo(denom=...)` so that the mean over ranks
    of the local losses — what DDP's gradient averaging computes — equals the single-process loss of vrnn.py:277."""
    total = torch.as_tensor(x_sl_local).sum().to(torch.float64).reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        dev_total = total.cuda() if backend == "nccl" else total
        dist.all_reduce(dev_total, op=dist.ReduceOp.SUM, group=group)
        return float(dev_total.item()) / dist.get_world_size(group)
    return float(total.item())


class _Pending:
    """Handle of an in-flight sums all-reduce: `.wait()` makes the current stream wait for it and returns the tensor."""

    def __init__(self, tensor, work):
        self.tensor, self.work = tensor, work

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        return self.tensor


def all_reduce_sums(sums: torch.Tensor, group=None, async_op: bool = False, inplace: bool = False):
    """All-reduce(sum) of a `fused_elbo` result's `sums` (8,) over the ranks — the path's only exchange.  Every entry
    is summed; the ratio entries (loss, bits-per-dim, nansum-loss) must afterwards be recomputed with `combine_sums`.
    `async_op=True` returns a handle (`.wait()`), so the collective (NCCL runs it on its own stream) overlaps the next
    step's kernels; `inplace=True` reduces into `sums` itself instead of a copy."""
    out = sums.detach() if inplace else sums.detach().clone()
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if async_op:
        return _Pending(out, work)
    return out


def combine_sums(sums: torch.Tensor, beta: float) -> torch.Tensor:
    """Recompute the ratio entries (loss, bits-per-dim; indices as in blvm_elbo_finalize) from globally reduced additive
    sums [.., sum log_prob, sum kl, sum kl_fn, sum elbo, sum x_sl, ..]; the nansum-loss entry is set to the loss."""
    out = sums.clone()
    s_logp, s_kl, s_klfn, s_elbo, s_len = out[1], out[2], out[3], out[4], out[5]
    out[0] = -(s_logp - beta * s_klfn) / s_len
    out[6] = -s_elbo / 0.6931471805599453 / s_len
    out[7] = out[0]
    return out

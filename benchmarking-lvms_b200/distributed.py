"""Batch-sharded data parallelism for the path: utterances are independent, so each rank runs the
kernels on its own B/W rows; the only exchange is one all-reduce of the scalar sums (NCCL over NVLink on GPUs, gloo in
the CPU tests).  No collective touches per-sample data."""
from typing import Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_rows", "global_denominator", "all_reduce_sums", "combine_sums", "SumsExchange", "bind_to_gpu_numa_node"]


def bind_to_gpu_numa_node(device_index: int):
    """One process per GPU: run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE allocating pinned host
    buffers (first touch places them on that node).  Otherwise about half the ranks of an 8-GPU box stage their
    host->device copies through the socket interconnect, which caps the aggregate H2D rate well below 8 PCIe links.
    Best effort: returns the node number, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        else:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device_index
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else str(bus)).lower()
            if len(bus.split(":")[0]) == 8:        # NVML prints an 8-digit domain, sysfs uses 4
                bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


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


class SumsExchange:
    """The scalar-sum exchange fused into the finalize kernel over NVLink peer memory (single node, <= 8 ranks).

    Each rank owns a small symmetric buffer (torch symmetric memory: one allocation mapped into every rank).  With
    `fused_elbo(..., exchange=ex)` the last CTA of the finalize kernel stores the step's 8 fp64 sums into this rank's
    slot of EVERY rank's buffer and releases a flag — no NCCL call, no extra launch, nothing on the host.  The same
    kernel also adds up (in rank order: bit-reproducible) the slots all ranks published for the PREVIOUS step into
    `self.global_sums` (8,) = [global loss, sum log_prob, sum kl, sum kl_fn, sum elbo, sum x_sl, global bpd, step]: they
    landed a whole step ago, so this never stalls, and it bounds the skew between ranks, which makes the 4-deep slot ring
    safe.  `consume(lag=0)` (a one-warp kernel) fetches the sums of the latest step, e.g. after the last step of an
    epoch.  All state is on the device: the calls can be captured in a CUDA graph.

    Contract: every rank issues the SAME number of steps with `exchange=ex` (like any collective).  A rank that runs ahead of a peer
    that has stopped waits in its finalize kernel for that peer's previous-step slot, gives up after ~10 s, sets bit 0 of `self.err`
    and leaves `global_sums` incomplete; nothing on the host notices by itself, so call `check()` (one sync) wherever the global
    sums are read on the host, e.g. once per logging interval / at the end of an epoch (bench.py does, after the timed region).
    With an uneven last batch, run that step without `exchange=` and use `all_reduce_sums` instead.
    """

    def __init__(self, group=None, device=None):
        import torch.distributed._symmetric_memory as symm

        from ._lib import lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("SumsExchange is a single-node exchange (<= 8 ranks); use all_reduce_sums across nodes")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n = int(lib.blvm_exchange_buffer_bytes()) // 8
        self.buffer = symm.empty(n, dtype=torch.float64, device=self.device)
        self.buffer.zero_()
        self.handle = symm.rendezvous(self.buffer, self.group)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.counters = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.global_sums = torch.zeros(8, dtype=torch.float64, device=self.device)   # previous step's global sums
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        self.handle.barrier()              # every buffer is zeroed before anyone publishes
        torch.cuda.synchronize(self.device)

    def consume(self, beta: float = 1.0, lag: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Global sums of step (published - lag); `out[7]` tells which step they belong to (0 = nothing consumed yet)."""
        from . import ops
        from ._lib import check, lib
        if out is None:
            out = torch.zeros(8, dtype=torch.float64, device=self.device)
        with ops._on_device(self.device):
            rc = lib.blvm_exchange_consume(self.buffer.data_ptr(), self.world, self.counters.data_ptr(), int(lag), float(beta),
                                           out.data_ptr(), self.err.data_ptr(), ops._stream(self.device.index))
            check(rc, "blvm_exchange_consume")
        ops._count()
        return out

    def check(self):
        """One sync: raise if a consume timed out (a peer never published) or a slot was overrun."""
        e = int(self.err.item())
        if e:
            self.err.zero_()
            raise RuntimeError(f"blvm_b200 SumsExchange error flags {e:#x} (1 = timeout waiting for a peer, 2 = slot overrun)")

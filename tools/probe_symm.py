"""Probe: does torch symmetric memory (peer-mapped buffers over NVLink) work on this box?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm.empty(1024, dtype=torch.float64, device=f"cuda:{local}")
t.zero_()
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None, flush=True)
h.barrier()
peer = (rank + 1) % world
remote = h.get_buffer(peer, (8,), torch.float64, rank * 8)
remote.fill_(float(rank + 1))
h.barrier()
torch.cuda.synchronize()
print(rank, "local buffer head", t[: 8 * world : 8].tolist(), flush=True)
dist.barrier()
dist.destroy_process_group()

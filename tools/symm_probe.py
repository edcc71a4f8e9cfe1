"""Probe: does torch symmetric memory rendezvous work on this box (needed for the fused peer-store gather)?"""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(world * 4, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], flush=True)
t.fill_(-1.0)
hdl.barrier()
# every rank writes its slice into every peer's buffer through the peer pointers (plain torch copies here)
for r in range(world):
    peer = hdl.get_buffer(r, (world * 4,), torch.float32)
    peer[rank * 4:(rank + 1) * 4] = float(rank)
torch.cuda.synchronize()
hdl.barrier()
torch.cuda.synchronize()
print(rank, t.tolist(), flush=True)
dist.destroy_process_group()

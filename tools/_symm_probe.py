import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = sm.empty(1024, dtype=torch.float32, device=torch.device("cuda", local))
t.fill_(rank + 1)
h = sm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "sig", [hex(p) for p in h.signal_pad_ptrs], "pad size", h.signal_pad_size, "mc", getattr(h, "multicast_ptr", None), flush=True)
h.barrier()
peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
h.barrier()
dist.destroy_process_group()

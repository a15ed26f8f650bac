"""Developer probe: does torch symmetric memory give peer pointers a kernel can store through on this box?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
n = 1 << 24
t = symm.empty(n, dtype=torch.float64, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
t.fill_(-1.0)
hdl.barrier()
peer = (rank + 1) % world
remote = hdl.get_buffer(peer, (n,), torch.float64)
src = torch.full((n,), float(rank), dtype=torch.float64, device=dev)
torch.cuda.synchronize(); hdl.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    a.record(); remote.copy_(src); b.record(); torch.cuda.synchronize()
    print(rank, "peer write", n * 8 / a.elapsed_time(b) / 1e6, "GB/s", flush=True)
hdl.barrier()
torch.cuda.synchronize()
print(rank, "got", float(t[0]), float(t[-1]), "expect", float((rank - 1) % world), flush=True)
# sub-group rendezvous
if world >= 4:
    G = world // 2
    groups = [dist.new_group(ranks=list(range(i * G, (i + 1) * G))) for i in range(2)]
    pg = groups[rank // G]
    t2 = symm.empty(1024, dtype=torch.float64, device=dev)
    h2 = symm.rendezvous(t2, pg)
    print(rank, "group ptrs", len(h2.buffer_ptrs), h2.rank, h2.world_size, flush=True)
dist.barrier(); dist.destroy_process_group()

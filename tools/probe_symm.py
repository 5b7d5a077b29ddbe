"""Probe: torch symmetric memory on this box (peer-mapped buffers for the fused dW2 reduce-scatter)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1024, 384, dtype=torch.float32, device=dev)
t.zero_()
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "world", hdl.world_size, flush=True)
dist.barrier(); torch.cuda.synchronize()
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (1024, 384), torch.float32)
pt.add_(float(rank + 1))          # plain ATen kernel writing PEER memory through the mapping
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
print(rank, "my buffer after peer add:", t[0, 0].item(), "(expected", float(((rank - 1) % world) + 1), ")", flush=True)
dist.destroy_process_group()

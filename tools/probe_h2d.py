"""H2D bandwidth of one pinned 103 MB blob per rank: alone vs all ranks at once, with and without binding the
rank to the CPUs of its GPU's NUMA node before the pinned allocation (first touch).  Run under torchrun."""
import os, sys, glob, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
def rd(p):
    try: return open(p).read().strip()
    except Exception as e: return f"? ({type(e).__name__})"
bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(local)
busid = pynvml.nvmlDeviceGetPciInfo(h).busId
busid = busid.decode() if isinstance(busid, bytes) else busid
short = busid[-12:].lower()
node = rd(f"/sys/bus/pci/devices/{short}/numa_node")
nodes = {os.path.basename(p): rd(p + "/cpulist") for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}
aff = sorted(os.sched_getaffinity(0))
if rank == 0:
    print("nodes", nodes, "cpu_count", os.cpu_count(), flush=True)
print(f"rank {rank} gpu {busid} numa_node {node} affinity {aff[:4]}..{aff[-2:]} ({len(aff)})", flush=True)
N = 103 * 1024 * 1024
def bw(blob, reps=20, sync_all=False):
    d = torch.empty(N, dtype=torch.uint8, device=dev)
    for _ in range(3): d.copy_(blob, non_blocking=True)
    torch.cuda.synchronize()
    if sync_all and world > 1: dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): d.copy_(blob, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return N * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
def parse(cl):
    out = []
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part.strip().isdigit(): out.append(int(part))
    return out
blob0 = torch.empty(N, dtype=torch.uint8).pin_memory(); blob0.fill_(1)
res = {}
for r in range(world):          # alone
    if world > 1: dist.barrier()
    if r == rank: res["alone"] = bw(blob0)
    if world > 1: dist.barrier()
res["all"] = bw(blob0, sync_all=True)
bound = "n/a"
try:
    cpus = [c for c in parse(nodes.get(f"node{node}", "")) if c in aff] if node not in ("-1", "") and not node.startswith("?") else []
    if cpus:
        os.sched_setaffinity(0, cpus); bound = f"{len(cpus)} cpus of node{node}"
    else:
        bound = "no local cpus allowed"
except Exception as e:
    bound = f"failed {e}"
blob1 = torch.empty(N, dtype=torch.uint8).pin_memory(); blob1.fill_(2)
res["all_bound"] = bw(blob1, sync_all=True)
print(f"rank {rank} H2D GB/s alone {res['alone']:.1f} all {res['all']:.1f} all_bound {res['all_bound']:.1f} ({bound})", flush=True)
if world > 1:
    dist.barrier(); torch.cuda.synchronize(); os._exit(0)

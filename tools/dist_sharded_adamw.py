"""Sharded optimizer step on real GPUs over NCCL (second half): run under
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_sharded_adamw.py
Every rank starts from the same head parameters and has its own gradients; after each step
  * all replicas are bit-identical,
  * they match torch.optim.AdamW applied to the rank-mean gradient (<= 1e-6 relative),
  * `last_grad_norm` is the norm of the rank-mean gradient,
  * `consolidated_state_dict()` matches torch's moments.
Prints PASS / FAIL per check and the time of the step against a full (unsharded) FusedAdamW step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dinox_b200.optim import FusedAdamW, ShardedFusedAdamW

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
D, K = 384, 65536
shapes = [(D, D), (D,), (K, D), (K,)]
g0 = torch.Generator().manual_seed(5)
init = [torch.randn(s, generator=g0) * 0.05 for s in shapes]
params = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
ref = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
hp = dict(lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.04)
opt = ShardedFusedAdamW(params, process_group=None, **hp)
topt = torch.optim.AdamW(ref, **hp)
assert [s for s, _ in opt.layout] == [False, False, True, False], opt.layout
ok = True


def check(name, cond, detail=""):
    global ok
    ok = ok and bool(cond)
    if rank == 0:
        print(f"[{name}] {'PASS' if cond else 'FAIL'} {detail}", flush=True)


for it in range(5):
    grads = []
    for r in range(world):
        gr = torch.Generator().manual_seed(1000 * it + r)
        grads.append([torch.randn(s, generator=gr) * (0.01 + 0.02 * r) for s in shapes])
    for p, g in zip(params, grads[rank]):
        p.grad = g.to(dev)
    mean = [torch.stack([grads[r][i] for r in range(world)]).to(dev).mean(0) for i in range(len(shapes))]
    for p, g in zip(ref, mean):
        p.grad = g
    opt.step()
    topt.step()
    torch.cuda.synchronize()
    err = max(float((a.data - b.data).abs().max() / b.data.abs().max()) for a, b in zip(params, ref))
    check(f"step {it} vs torch.optim.AdamW on the mean gradient", err < 1e-6, f"max rel err {err:.2e}")
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in mean)))
    check(f"step {it} grad norm", abs(float(opt.last_grad_norm) - gn) / gn < 1e-5, f"{float(opt.last_grad_norm):.6f} vs {gn:.6f}")
    flat = torch.cat([p.data.view(-1) for p in params])
    allf = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(allf, flat)
    check(f"step {it} replicas bit-identical", all(torch.equal(allf[0], x) for x in allf))
sd, tsd = opt.consolidated_state_dict(), topt.state_dict()
e = max(float((sd["state"][i][k] - tsd["state"][i][k]).abs().max() / (tsd["state"][i][k].abs().max() + 1e-30))
        for i in range(len(shapes)) for k in ("exp_avg", "exp_avg_sq"))
check("consolidated moments vs torch", e < 1e-5, f"max rel err {e:.2e}")


def timeit(fn, n=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


full = FusedAdamW([torch.nn.Parameter(t.clone().to(dev)) for t in init], **hp)
for p in full.param_groups[0]["params"]:
    p.grad = torch.randn_like(p) * 0.01


def full_step():   # what a DDP wrapper + replicated optimizer does: all-reduce every gradient, full AdamW
    for p in full.param_groups[0]["params"]:
        dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
    full.step()


t_sh, t_full = timeit(opt.step), timeit(full_step)
if rank == 0:
    print(f"step time, {world} ranks, 25.3 M head parameters: sharded {t_sh:.3f} ms, all-reduce + replicated {t_full:.3f} ms", flush=True)
    print("SHARDED ADAMW", "PASS" if ok else "FAIL", flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0 if ok else 1)

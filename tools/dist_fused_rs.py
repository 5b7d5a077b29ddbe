"""Fused dW2 GEMM + reduce-scatter over peer memory (second half) against the NCCL path.  Run under
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_fused_rs.py [time]
Checks (every rank its own crops, C2-like shapes at reduced size unless `time`):
  1. gradient: the peer-mapped shard after one backward == NCCL reduce_scatter(AVG) of the locally accumulated dW2
  2. optimizer: ShardedFusedAdamW(grad_shards=...) over two accumulation windows == ShardedFusedAdamW with NCCL
     reduce-scatter (parameters agree to fp32 summation-order noise), replicas identical
  3. `time`: one accumulation window (4 micro-steps + optimizer step) at C2 shapes, both variants, CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from dinox_b200 import losshead, synth
from dinox_b200.optim import PeerGradShards, ShardedFusedAdamW
from dinox_b200.step import LossHeadStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
timing = len(sys.argv) > 1 and sys.argv[1] == "time"
cfg = dict(synth.CONFIGS["C2"]) if timing else dict(batch=8, dim=384, out_dim=8192, n_patches=36)
sh = synth.LossHeadShapes(**cfg)
ACC = 4


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def make(fused: bool):
    st = LossHeadStep(sh, dev, accum=ACC, process_group=True, with_backbone_params=False)
    head = list(st.student_head.parameters())
    shards = {head[2]: PeerGradShards(head[2], True)} if fused else None
    st.w2_grad_shards = shards[head[2]] if fused else None
    opt = ShardedFusedAdamW(head, lr=1e-3, weight_decay=0.04, process_group=True, grad_shards=shards)
    return st, opt, head


def batches(seed):
    """this rank's ACC micro-batches of a window, resident on the device"""
    out = []
    for m in range(ACC):
        f = synth.feature_batch(sh, synth.seeded_generator(seed * 10 + m, rank), patches_from_tokens=True)
        out.append({k: v.to(dev) for k, v in f.items()})
    return out


def window(st, opt, feats):
    """one accumulation window on this rank's crops: ACC micro-steps (fwd + bwd), optimizer step"""
    for i, f in enumerate(feats):
        if st.w2_grad_shards is not None:
            st.w2_grad_shards.flush = i == len(feats) - 1     # the window's gradient leaves with the last micro-step
        fd = {k: (v.detach().requires_grad_(True) if k.startswith("student") else v) for k, v in f.items()}
        out, loss = st._losses(fd)
        loss.backward()
    opt.step()
    opt.zero_grad()
    losshead.invalidate_weight_cache()


ok = True
# ---- 1. gradient of one backward
st_a, opt_a, head_a = make(False)
st_b, opt_b, head_b = make(True)
f = synth.feature_batch(sh, synth.seeded_generator(5, rank), patches_from_tokens=True)
for st in (st_a, st_b):
    fd = {k: (v.to(dev).requires_grad_(True) if k.startswith("student") else v.to(dev)) for k, v in f.items()}
    out, loss = st._losses(fd)
    loss.backward()
rows = head_a[2].shape[0] // world
ref = torch.empty(rows, head_a[2].shape[1], device=dev)
dist.reduce_scatter_tensor(ref.view(-1), head_a[2].grad.view(-1), op=dist.ReduceOp.AVG)
dist.barrier()
torch.cuda.synchronize()
e = rel(st_b.w2_grad_shards.acc, ref)
t = torch.tensor([e], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    good = t.item() < 2e-6
    ok &= good
    print(f"[fused reduce-scatter] world {world}: peer-mapped dW2 shard vs NCCL reduce_scatter(AVG): max rel err over ranks "
          f"{t.item():.2e} (limit 2e-06) {'PASS' if good else 'FAIL'}", flush=True)
assert head_b[2].grad is None, "the fused path must not materialise the full dW2"
for p in head_a + head_b:
    p.grad = None
st_b.w2_grad_shards.zero_()
dist.barrier()

# ---- 2. optimizer over two windows
for w in range(2):
    fb = batches(20 + w)
    window(st_a, opt_a, fb)
    window(st_b, opt_b, fb)
torch.cuda.synchronize()
errs = [rel(b, a) for a, b in zip(head_a, head_b)]
t = torch.tensor(errs, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
w2 = head_b[2].detach().clone()
w2_0 = w2.clone()
dist.broadcast(w2_0, 0)
same = torch.tensor([float(torch.equal(w2, w2_0))], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    good = max(t.tolist()) < 5e-6 and same.item() == 1.0
    ok &= good
    print(f"[fused reduce-scatter] parameters after 2 windows vs the NCCL path: rel err W1 {t[0].item():.1e} b1 {t[1].item():.1e} "
          f"W2 {t[2].item():.1e} b2 {t[3].item():.1e} (limit 5e-06); replicas identical: {bool(same.item())} "
          f"{'PASS' if good else 'FAIL'}", flush=True)

# ---- 3. timing of a window
if timing:
    fb = batches(40)

    import statistics

    def one(st, opt):
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        window(st, opt, fb)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    for _ in range(2):
        one(st_a, opt_a); one(st_b, opt_b)
    ta, tb = [], []
    for _ in range(12):                 # interleaved A/B: host jitter of the eager launches hits both alike
        ta.append(one(st_a, opt_a))
        tb.append(one(st_b, opt_b))
    ms_nccl, ms_fused = statistics.median(ta), statistics.median(tb)
    if rank == 0:
        print(f"[fused reduce-scatter] C2 window (4 eager micro-steps + sharded AdamW), world {world}, max over ranks, median of 12 "
              f"interleaved windows: NCCL reduce-scatter {ms_nccl:.3f} ms (min {min(ta):.3f}), fused GEMM+reduce-scatter "
              f"{ms_fused:.3f} ms (min {min(tb):.3f}): {ms_nccl - ms_fused:+.3f} ms per window", flush=True)

dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("FUSED RS", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()

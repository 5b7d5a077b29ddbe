"""Data-parallel parity on real GPUs over NCCL.  Run under
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_parity_nccl.py
(tests/test_gpu_dist.py does, when the box has >= 2 GPUs).

Every rank owns B images (all their views / masked tokens).  With process_group=True the sharded fused
loss head must reproduce the single-process result on the concatenated batch:
  * losses: mean over ranks of the rank-local means == loss of the full batch (centre teacher: per-sample
    terms; Sinkhorn-Knopp: per-prototype sums all-gathered in the log domain),
  * centre / patch centre after the update (all-reduced sums of the teacher activations),
  * head gradients: the fused loss inside DistributedDataParallel (FusedLossHead, autograd-visible head
    gradients) leaves every rank with the gradient of the full-batch loss (DDP averages),
  * ranks that mask different numbers of tokens still produce the full-batch patch centre (the mean divides
    by the all-reduced row count on the device).
Rank 0 evaluates the full batch alone as the reference and prints one line per check and a final verdict."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from dinox_b200 import losshead, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, Vg, Vl, D, K, n_mask = 8, 2, 2, 128, 4096, 5           # B images per rank
V = Vg + Vl
gen = torch.Generator().manual_seed(77)
Bg = B * world
sd_s, sd_t = synth.head_weights(D, K, gen), synth.head_weights(D, K, gen)
full = dict(student_cls=torch.randn(V, Bg, D, generator=gen), teacher_cls=torch.randn(Vg, Bg, D, generator=gen),
            student_patch=torch.randn(Vg, Bg, n_mask, D, generator=gen),
            teacher_patch=torch.randn(Vg, Bg, n_mask, D, generator=gen))
c0 = torch.randn(1, K, generator=gen) * 0.05


def shard(lo, hi):
    """view-major rows of images [lo, hi)"""
    n = hi - lo
    return dict(student_cls=full["student_cls"][:, lo:hi].reshape(V * n, D),
                teacher_cls=full["teacher_cls"][:, lo:hi].reshape(Vg * n, D),
                student_patch=full["student_patch"][:, lo:hi].reshape(Vg * n * n_mask, D),
                teacher_patch=full["teacher_patch"][:, lo:hi].reshape(Vg * n * n_mask, D),
                masks_weight=torch.full((Vg * n * n_mask,), 1.0 / n_mask))


def build(mode, pg):
    s_head, t_head = losshead.ProjectionHead(D, K), losshead.ProjectionHead(D, K)
    s_head.load_state_dict(sd_s)
    t_head.load_state_dict(sd_t)
    dl = losshead.DINOLoss(K, 0.9, n_global=Vg, n_local=Vl, teacher_mode=mode, process_group=pg)
    mod = losshead.FusedLossHead(s_head, t_head, dl).to(dev)
    mod.dino_loss.center.copy_(c0)
    mod.center_patch.copy_(c0)
    return mod


def run(f, mode, pg, ddp):
    mod = build(mode, pg)
    net = torch.nn.parallel.DistributedDataParallel(mod, device_ids=[local]) if ddp else mod
    fd = {k: v.to(dev) for k, v in f.items()}
    loss = net(fd["student_cls"], fd["teacher_cls"], 0.1, 0.04, student_patch=fd["student_patch"],
               teacher_patch=fd["teacher_patch"], masks_weight=fd["masks_weight"])
    loss.backward()
    grads = [p.grad.detach().clone() for p in mod.student_head.parameters()]
    return mod.last["loss_dino"], mod.last["loss_ibot"], mod.dino_loss.center.clone(), mod.center_patch.clone(), grads


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


ok = True
for mode in ("center", "sinkhorn"):
    ld, li, c, cp, g = run(shard(rank * B, (rank + 1) * B), mode, True, ddp=True)
    t = torch.stack([ld, li])
    dist.all_reduce(t)
    t /= world                                              # mean of the rank-local means
    if rank == 0:
        ld1, li1, c1, cp1, g1 = run(shard(0, Bg), mode, None, ddp=False)
        checks = {"loss_dino": (abs(t[0].item() - ld1.item()) / abs(ld1.item()), 2e-5),
                  "loss_ibot": (abs(t[1].item() - li1.item()) / abs(li1.item()), 2e-5),
                  "center": (rel(c, c1), 2e-5), "center_patch": (rel(cp, cp1), 2e-5)}
        # gradients pass bf16 backward GEMMs whose operand (dL/dlogits) is rounded per rank vs per full batch
        for n, a, b in zip(("dW1", "db1", "dW2", "db2"), g, g1):
            checks["ddp_" + n] = (rel(a, b), 4e-3)
        print(f"[{mode}] world {world}: sharded mean loss_dino {t[0].item():.9f} ibot {t[1].item():.9f} | "
              f"full batch {ld1.item():.9f} {li1.item():.9f} | rank0 local {ld.item():.9f}", flush=True)
        for k, (v, lim) in checks.items():
            good = v < lim
            ok &= good
            print(f"[{mode}] {k}: rel err {v:.2e} (limit {lim:.0e}) {'PASS' if good else 'FAIL'}", flush=True)

# ranks may mask different numbers of tokens: the patch-centre mean divides by the all-reduced row count
def truncated(r):
    f = shard(r * B, (r + 1) * B)
    drop = 2 * r                                            # rank r masks 2r tokens fewer
    if drop:
        for k in ("student_patch", "teacher_patch", "masks_weight"):
            f[k] = f[k][:-drop]
    return f


mod = build("center", True)
fd = {k: v.to(dev) for k, v in truncated(rank).items()}
mod(fd["student_cls"], fd["teacher_cls"], 0.1, 0.04, student_patch=fd["student_patch"], teacher_patch=fd["teacher_patch"],
    masks_weight=fd["masks_weight"])
if rank == 0:
    parts = [truncated(r) for r in range(world)]
    f1 = shard(0, Bg)
    for k in ("student_patch", "teacher_patch", "masks_weight"):
        f1[k] = torch.cat([p[k] for p in parts], 0)
    ref = build("center", None)
    f1 = {k: v.to(dev) for k, v in f1.items()}
    ref(f1["student_cls"], f1["teacher_cls"], 0.1, 0.04, student_patch=f1["student_patch"], teacher_patch=f1["teacher_patch"],
        masks_weight=f1["masks_weight"])
    for name, a, b in (("center", mod.dino_loss.center, ref.dino_loss.center), ("center_patch", mod.center_patch, ref.center_patch)):
        v = rel(a, b)
        good = v < 2e-5
        ok &= good
        print(f"[unequal masked rows] {name}: rel err {v:.2e} (limit 2e-05) {'PASS' if good else 'FAIL'}", flush=True)

dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("DIST PARITY", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()

"""Data-parallel parity on real GPUs over NCCL: run under
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_parity_nccl.py
Every rank owns half of the images (all their views / masked tokens); the sharded fused loss head with
process_group=True must reproduce the single-process result on the concatenated batch:
  * centre / patch-centre after the update (all-reduced sum of teacher activations),
  * Sinkhorn-Knopp CLS loss (per-prototype sums all-gathered in the log domain): mean of the rank losses
    == single-process loss on the full batch,
  * centre-mode losses are per-sample, so the same identity holds there too.
Rank 0 also evaluates the full batch alone and prints PASS / FAIL per check."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import dinox_b200 as dx
from dinox_b200 import synth

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
            student_patch=torch.randn(Vg, Bg, n_mask, D, generator=gen), teacher_patch=torch.randn(Vg, Bg, n_mask, D, generator=gen))
c0 = torch.randn(1, K, generator=gen) * 0.05


def shard(lo, hi):
    """view-major rows of images [lo, hi)"""
    n = hi - lo
    return dict(student_cls=full["student_cls"][:, lo:hi].reshape(V * n, D), teacher_cls=full["teacher_cls"][:, lo:hi].reshape(Vg * n, D),
                student_patch=full["student_patch"][:, lo:hi].reshape(Vg * n * n_mask, D),
                teacher_patch=full["teacher_patch"][:, lo:hi].reshape(Vg * n * n_mask, D),
                masks_weight=torch.full((Vg * n * n_mask,), 1.0 / n_mask))


def run(f, mode, pg):
    s_head, t_head = dx.DinoStudentTeacher.__mro__ and None, None
    from dinox_b200 import losshead
    s_head, t_head = losshead.ProjectionHead(D, K).to(dev), losshead.ProjectionHead(D, K).to(dev)
    s_head.load_state_dict(sd_s); t_head.load_state_dict(sd_t)
    dl = dx.DINOLoss(K, 0.9, n_global=Vg, n_local=Vl, teacher_mode=mode, process_group=pg).to(dev)
    dl.center.copy_(c0)
    cp = c0.clone().to(dev)
    fd = {k: v.to(dev) for k, v in f.items()}
    out = dx.fused_head_dino_loss(fd["student_cls"], fd["teacher_cls"], s_head, t_head, dl, 0.1, 0.04,
                                  student_patch=fd["student_patch"], teacher_patch=fd["teacher_patch"],
                                  masks_weight=fd["masks_weight"], center_patch=cp)
    return out["loss_dino"].detach(), out["loss_ibot"].detach(), dl.center.clone(), cp


ok = True
for mode in ("center", "sinkhorn"):
    ld, li, c, cp = run(shard(rank * B, (rank + 1) * B), mode, True)
    t = torch.stack([ld, li]); dist.all_reduce(t); t /= world          # mean of the rank-local means
    if rank == 0:
        ld1, li1, c1, cp1 = run(shard(0, Bg), mode, None)
        rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        checks = {"loss_dino": abs(t[0].item() - ld1.item()) / abs(ld1.item()), "loss_ibot": abs(t[1].item() - li1.item()) / abs(li1.item()),
                  "center": rel(c, c1), "center_patch": rel(cp, cp1)}
        print(f"[{mode}] sharded mean loss_dino {t[0].item():.9f} ibot {t[1].item():.9f} | full batch {ld1.item():.9f} {li1.item():.9f} | rank0 local {ld.item():.9f}", flush=True)
        for k, v in checks.items():
            good = v < 2e-5
            ok &= good
            print(f"[{mode}] {k}: rel err {v:.2e} {chr(80)+chr(65)+chr(83)+chr(83) if good else chr(70)+chr(65)+chr(73)+chr(76)}", flush=True)
dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("DIST PARITY", "PASS" if ok else "FAIL", flush=True)
os._exit(0)

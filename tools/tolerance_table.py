"""Tolerance table: error of THIS library against the fp32 reference arithmetic, beside the error of
the as-run reference under CUDA bf16 autocast against the same fp32 truth - so a tolerance in the tests can be read
as "autocast-equivalent" or not.  Run on a B200:  python tools/tolerance_table.py > profiles/r02_tolerance_table.txt

  truth     : oracle (the reference's arithmetic, restated) in fp32 on the CPU
  autocast  : the same oracle code on CUDA inside torch.autocast("cuda", dtype=torch.bfloat16) - what
              scripts/phase5_big_run.py --amp executes (Linear / bmm in bf16, softmax in fp32)
  ours      : dinox_b200 (bf16 tensor-core operands, fp32 accumulation, logits / Gram never rounded)
Cases: (1) the reference micro-step golden (features produced by the reference's PatchViT on synthetic CT crops,
2 global views, K=256, D=32); (2) C1 at full size (K=65536, 80/16/928 rows, multi-crop + iBOT + Gram)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dinox_b200 import losshead as dx, synth
from dinox_b200.step import LossHeadStep
from oracle import losshead_oracle as O

DEV = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def oracle_run(head_s, head_t, f, K, mode, Vg, Vl, c0, cp0, device, autocast):
    sp = O.HeadParams(*[p.clone().to(device).requires_grad_(True) for p in head_s])
    tp = O.HeadParams(*[p.clone().to(device) for p in head_t])
    orc = O.LossHeadOracle(sp, tp, K, center_momentum=0.9, n_global=Vg, n_local=Vl, teacher_mode=mode, policy="fp32",
                           patch_teacher_mode="center")
    orc.center, orc.center_patch = c0.clone().to(device), cp0.clone().to(device)
    fo = {k: (v.clone().to(device).requires_grad_(True) if k.startswith("student") else v.to(device)) for k, v in f.items()}
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
    with ctx:
        out = orc.step(fo["student_cls"], fo["teacher_cls"], 0.1, 0.04, student_tok=fo.get("student_tok"),
                       teacher_tok=fo.get("teacher_tok"), student_patch=fo.get("student_patch"),
                       teacher_patch=fo.get("teacher_patch"), masks_weight=fo.get("masks_weight"), accum=1)
    res = {k: v.float() for k, v in out.items() if k.startswith("loss_")}
    res["d_student_cls"] = fo["student_cls"].grad
    if "student_patch" in fo:
        res["d_student_patch"] = fo["student_patch"].grad
    if "student_tok" in fo:
        res["d_student_tok"] = fo["student_tok"].grad
    for n, p in zip(("dW1", "db1", "dW2", "db2"), sp.tensors()):
        res[n] = p.grad
    res["center"] = orc.center
    if "student_patch" in fo:
        res["center_patch"] = orc.center_patch
    return res


def ours_run(sh, head_s, head_t, f, mode, c0, cp0):
    st = LossHeadStep(sh, DEV, accum=1, with_backbone_params=False, teacher_mode=mode, center_momentum=0.9)
    with torch.no_grad():
        for p, q in zip(st.student_head.parameters(), head_s):
            p.copy_(q)
        for p, q in zip(st.teacher_head.parameters(), head_t):
            p.copy_(q)
    dx.invalidate_weight_cache()
    st.dino_loss.center.copy_(c0)
    st.center_patch.copy_(cp0)
    fd = {k: (v.to(DEV).requires_grad_(True) if k.startswith("student") else v.to(DEV)) for k, v in f.items()}
    out, loss = st._losses(fd)
    loss.backward()
    torch.cuda.synchronize()
    res = {k: v.float() for k, v in out.items() if k in ("loss_dino", "loss_ibot", "loss_gram")}
    res["d_student_cls"] = fd["student_cls"].grad
    if "student_patch" in fd:
        res["d_student_patch"] = fd["student_patch"].grad
    if "student_tok" in fd:
        res["d_student_tok"] = fd["student_tok"].grad
    for n, p in zip(("dW1", "db1", "dW2", "db2"), st.student_head.parameters()):
        res[n] = p.grad
    res["center"] = st.dino_loss.center
    if "student_patch" in fd:
        res["center_patch"] = st.center_patch
    return res


def table(title, truth, auto, ours):
    print(f"\n## {title}")
    print(f"{'quantity':18s} {'ours vs fp32':>14s} {'autocast vs fp32':>18s} {'ours vs autocast':>18s}")
    for k in truth:
        a, b, c = rel(ours[k], truth[k]), rel(auto[k], truth[k]), rel(ours[k], auto[k])
        print(f"{k:18s} {a:14.2e} {b:18.2e} {c:18.2e}")


def main():
    print("# tolerance table - relative L2 errors (scalars: relative error)")
    print(f"# torch {torch.__version__}, {torch.cuda.get_device_name(0)}; truth = oracle fp32 on CPU; autocast = oracle on CUDA under "
          "torch.autocast(bf16); ours = dinox_b200")
    # ---- (2) C1 at full size, centre and Sinkhorn-Knopp teachers
    sh = synth.LossHeadShapes(**synth.CONFIGS["C1"])
    for mode in ("center", "sinkhorn"):
        g = synth.seeded_generator(1, 0)
        f = synth.feature_batch(sh, g)
        c0 = torch.randn(1, sh.out_dim, generator=g) * 0.05
        cp0 = torch.randn(1, sh.out_dim, generator=g) * 0.05
        hs = list(synth.head_weights(sh.dim, sh.out_dim, g).values())
        ht = list(synth.head_weights(sh.dim, sh.out_dim, g).values())
        truth = oracle_run(hs, ht, f, sh.out_dim, mode, sh.n_global, sh.n_local, c0, cp0, "cpu", False)
        auto = oracle_run(hs, ht, f, sh.out_dim, mode, sh.n_global, sh.n_local, c0, cp0, DEV, True)
        ours = ours_run(sh, hs, ht, f, mode, c0, cp0)
        if mode == "sinkhorn":
            for r in (truth, auto, ours):
                r.pop("center", None)
        table(f"C1 full size (K=65536, D=384, 80 student / 16 teacher / 928 masked rows, Gram 200 tokens), teacher {mode}", truth, auto, ours)
    # ---- (1) reference micro-step golden: 2 global views, reference-shaped (no iBOT, no local crops)
    gp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "microstep.npz")
    g = np.load(gp)
    T = torch.from_numpy
    K, D = 256, 32
    sf, tf = T(g["student_feats"]), T(g["teacher_feats"])
    hs = [T(g["s_head_" + k]) for k in ("0_weight", "0_bias", "2_weight", "2_bias")]
    ht = [T(g["t_head_" + k]) for k in ("0_weight", "0_bias", "2_weight", "2_bias")]
    f = dict(student_cls=sf[:, 0].contiguous(), teacher_cls=tf[:, 0].contiguous(), student_tok=sf, teacher_tok=tf)
    c0 = T(g["center0"])
    B = sf.shape[0] // 2
    shm = synth.LossHeadShapes(batch=B, dim=D, out_dim=K, n_patches=sf.shape[1] - 5, n_global=2, n_local=0, mask_ratio=0.0)

    def fix(r):   # the golden differentiates w.r.t. the whole feature tensor: CLS gradient lives in row 0
        r = dict(r)
        dt = r.pop("d_student_tok").clone()
        dt[:, 0] += r.pop("d_student_cls")
        r["d_student_feats"] = dt
        r["loss_dino"], r["loss_gram"] = r["loss_dino"], r["loss_gram"]
        return r
    truth = fix(oracle_run(hs, ht, f, K, "center", 2, 0, c0, c0 * 0, "cpu", False))
    auto = fix(oracle_run(hs, ht, f, K, "center", 2, 0, c0, c0 * 0, DEV, True))
    ours = fix(ours_run(shm, hs, ht, f, "center", c0, c0 * 0))
    # the committed golden itself (produced by the reference's own classes) against the oracle truth
    gold = {"loss_dino": T(g["loss_dino"]), "loss_gram": T(g["loss_gram"]), "d_student_feats": T(g["d_student_feats"]),
            "center": T(g["center1"]), "dW1": T(g["g_head_0_weight"]), "db1": T(g["g_head_0_bias"]),
            "dW2": T(g["g_head_2_weight"]), "db2": T(g["g_head_2_bias"])}
    table("reference micro-step golden (tests/golden/microstep.npz: K=256, D=32, 2 views, reference PatchViT features)", truth, auto, ours)
    print("\n# oracle truth vs the committed golden (reference's own classes; accum scaling removed):")
    accum = float(g["accum"])
    for k, v in gold.items():
        scale = accum if k.startswith("d") else 1.0
        print(f"{k:18s} {rel(truth[k], v * scale):14.2e}")


if __name__ == "__main__":
    main()

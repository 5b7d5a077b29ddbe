"""Generate golden vectors for the loss-head path FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` is not present on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

It imports the reference's own classes (scripts/phase5_big_run.py: DINOLoss,
compute_gram_matrix, compute_gram_anchoring_loss, KoLeoLoss; zoo/arch.py:
DinoStudentTeacher, PatchViT), runs them on seeded inputs on CPU in fp32 and stores inputs and
outputs.  tests/test_oracle_golden.py then pins oracle/losshead_oracle.py against these files,
and the GPU parity tests compare the CUDA path against the same files.  The reference has no
golden vectors of its own for this path - these are the pin.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get("DINOX_REFERENCE", "/root/reference")
sys.path[:0] = [REF, os.path.join(REF, "scripts")]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def main():
    from phase5_big_run import DINOLoss, KoLeoLoss, compute_gram_anchoring_loss, compute_gram_matrix
    from zoo.arch import DinoStudentTeacher, PatchViT
    from dinox_b200 import synth

    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed reduction order

    # ---- KAT 1: uniform logits => ln K (docs/phase5_big_run.md:375-377)
    k = 8192
    l = DINOLoss(k, 0.9)
    kat1 = l(torch.zeros(4, k), torch.zeros(4, k), 0.1, 0.04)
    np.savez(os.path.join(OUT, "dino_uniform.npz"), out_dim=k, loss=_np(kat1), lnk=np.log(k))

    # ---- KAT 2: seeded (8,128), two consecutive calls
    g = torch.Generator().manual_seed(1234)
    s = torch.randn(8, 128, generator=g)
    t = torch.randn(8, 128, generator=g)
    l = DINOLoss(128, 0.9)
    s1 = s.clone().requires_grad_(True)
    loss_a = l(s1, t, 0.1, 0.04)
    loss_a.backward()
    center_a = l.center.clone()
    s2 = s.clone().requires_grad_(True)
    loss_b = l(s2, t, 0.1, 0.04)
    loss_b.backward()
    center_b = l.center.clone()
    # ---- KAT 3: Gram, continuing the same generator
    sf = torch.randn(4, 21, 32, generator=g)
    tf = torch.randn(4, 21, 32, generator=g)
    sfg = sf.clone().requires_grad_(True)
    gl = compute_gram_anchoring_loss(sfg, tf)
    gl.backward()
    gm = compute_gram_matrix(sf[:, 1:])
    # ---- KAT 4: KoLeo
    x = torch.randn(8, 128, generator=g)
    kl = KoLeoLoss()(x)
    np.savez(os.path.join(OUT, "dino_seeded.npz"),
             student=_np(s), teacher=_np(t), student_temp=0.1, teacher_temp=0.04, momentum=0.9,
             loss_a=_np(loss_a), grad_a=_np(s1.grad), center_a=_np(center_a),
             loss_b=_np(loss_b), grad_b=_np(s2.grad), center_b=_np(center_b),
             gram_student=_np(sf), gram_teacher=_np(tf), gram_loss=_np(gl), gram_grad=_np(sfg.grad),
             gram_matrix=_np(gm), gram_self=_np(compute_gram_anchoring_loss(tf, tf)),
             koleo_x=_np(x), koleo=_np(kl))

    # ---- KAT 5: larger / ragged DINOLoss shapes incl. non-multiple-of-4 K and big logits
    cases = {}
    for name, (b, kk, scale, mom, seed) in {"k1000": (6, 1000, 3.0, 0.999, 101), "k4099": (10, 4099, 0.5, 0.9, 102),
                                            "k65536": (4, 65536, 1.0, 0.9, 103)}.items():
        g = torch.Generator().manual_seed(seed)
        s = (torch.randn(2 * b, kk, generator=g) * scale)
        t = (torch.randn(2 * b, kk, generator=g) * scale)
        l = DINOLoss(kk, mom)
        l.center.copy_(torch.randn(1, kk, generator=g) * 0.1)
        c0 = l.center.clone()
        sg = s.clone().requires_grad_(True)
        loss = l(sg, t, 0.1, 0.04)
        loss.backward()
        if kk <= 4099:
            cases.update({f"{name}_student": _np(s), f"{name}_teacher": _np(t), f"{name}_grad": _np(sg.grad)})
        else:  # keep the file small: store the seed recipe + reductions only
            cases.update({f"{name}_grad_norm": _np(sg.grad.norm()), f"{name}_grad_head": _np(sg.grad[:, :64])})
        cases.update({f"{name}_shape": np.array([2 * b, kk]), f"{name}_scale": scale, f"{name}_mom": mom,
                      f"{name}_seed": seed,
                      f"{name}_center0": _np(c0) if kk <= 4099 else _np(c0[:, :64]),
                      f"{name}_loss": _np(loss),
                      f"{name}_center1": _np(l.center) if kk <= 4099 else _np(l.center[:, :64]),
                      f"{name}_center1_sum": _np(l.center.sum())})
    np.savez(os.path.join(OUT, "dino_shapes.npz"), **cases)

    # ---- KAT 6: projection head (zoo/arch.py:246-261) + EMA loop (:1798-1802) on a tiny model
    torch.manual_seed(42)
    bb_s = PatchViT(img_size=32, patch=16, dim=32, depth=1, heads=2, scale_aware=True)
    bb_t = PatchViT(img_size=32, patch=16, dim=32, depth=1, heads=2, scale_aware=True)
    student = DinoStudentTeacher(bb_s, out_dim=96)
    teacher = DinoStudentTeacher(bb_t, out_dim=96)
    xg = torch.Generator().manual_seed(5)
    cls = torch.randn(6, 32, generator=xg)
    head_out = student.head(cls)
    keys = list(student.state_dict().keys())
    head_sd = {k: _np(v) for k, v in student.head.state_dict().items()}
    ps_before = [_np(p) for p in student.parameters()]
    pt_before = [_np(p) for p in teacher.parameters()]
    ema = 0.996
    with torch.no_grad():
        for p_s, p_t in zip(student.parameters(), teacher.parameters()):
            p_t.data.mul_(ema).add_(p_s.data, alpha=1.0 - ema)
    pt_after = [_np(p) for p in teacher.parameters()]
    np.savez(os.path.join(OUT, "head_ema.npz"), cls=_np(cls), head_out=_np(head_out),
             head_keys=np.array([k for k in keys if k.startswith("head.")]),
             n_params=len(ps_before), ema=ema,
             **{f"head_{k.replace('.', '_')}": v for k, v in head_sd.items()},
             **{f"ps_{i}": v for i, v in enumerate(ps_before)},
             **{f"pt0_{i}": v for i, v in enumerate(pt_before)},
             **{f"pt1_{i}": v for i, v in enumerate(pt_after)})

    # ---- KAT 7: one whole micro-step of the reference loop (:1741-1772) on synthetic CT crops
    torch.manual_seed(7)
    bb_s = PatchViT(img_size=32, patch=8, dim=32, depth=1, heads=2, scale_aware=True)
    student = DinoStudentTeacher(bb_s, out_dim=256)
    bb_t = PatchViT(img_size=32, patch=8, dim=32, depth=1, heads=2, scale_aware=True)
    teacher = DinoStudentTeacher(bb_t, out_dim=256)
    teacher.load_state_dict(student.state_dict())
    with torch.no_grad():  # make teacher != student so the loss is not degenerate
        for p in teacher.parameters():
            p.add_(torch.randn(p.shape, generator=xg) * 0.02)
    for p in teacher.parameters():
        p.requires_grad_(False)
    g = synth.seeded_generator(cfg=1, rank=0)
    views, spacing = synth.multicrop_batch(4, g, n_global=2, n_local=0, global_size=32)
    batch = torch.cat(views, 0)
    spacing_2b = torch.cat([spacing, spacing], 0)
    dl = DINOLoss(256, 0.9)
    dl.center.copy_(torch.randn(1, 256, generator=xg) * 0.05)
    c0 = dl.center.clone()
    student_feats = student.backbone(batch, spacing=spacing_2b)
    student_feats.retain_grad()
    with torch.no_grad():
        teacher_feats = teacher.backbone(batch, spacing=spacing_2b)
    student_out = student.head(student_feats[:, 0])
    teacher_out = teacher.head(teacher_feats[:, 0])
    loss_dino = dl(student_out, teacher_out, 0.1, 0.04)
    loss_gram = compute_gram_anchoring_loss(student_feats, teacher_feats)
    accum = 4
    loss = (loss_dino + 1.0 * loss_gram) / accum
    loss.backward()
    np.savez(os.path.join(OUT, "microstep.npz"),
             crops=_np(batch), spacing=_np(spacing_2b),
             student_feats=_np(student_feats), teacher_feats=_np(teacher_feats),
             center0=_np(c0), center1=_np(dl.center), accum=accum,
             student_temp=0.1, teacher_temp=0.04, momentum=0.9,
             loss_dino=_np(loss_dino), loss_gram=_np(loss_gram),
             d_student_feats=_np(student_feats.grad),
             **{f"s_head_{k.replace('.', '_')}": _np(v) for k, v in student.head.state_dict().items()},
             **{f"t_head_{k.replace('.', '_')}": _np(v) for k, v in teacher.head.state_dict().items()},
             **{f"g_head_{n.replace('.', '_')}": _np(p.grad) for n, p in student.head.named_parameters()})
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f))} bytes")


def window_sequence_golden():
    """KAT 9: TWO optimizer windows of the reference loop on head features (scripts/phase5_big_run.py:1738-1802):
    accum = 2 micro-steps each, loss = (L_dino + 0.5 L_gram) / accum, backward, then torch.optim.AdamW on the student
    head (:1621) and the EMA of the teacher head (:1798-1802).  Pins the glue over time: the centre after every
    micro-step, the accumulated head gradients at each window end, the student head after AdamW and the teacher
    head after the EMA.  Inputs are fixed random features (no backbone), written next to the outputs."""
    from phase5_big_run import DINOLoss, compute_gram_anchoring_loss
    from zoo.arch import DinoStudentTeacher
    torch.manual_seed(11)
    torch.set_num_threads(1)

    class _BB(torch.nn.Module):          # DinoStudentTeacher only needs .dim of its backbone to build the head
        dim = 24
    D, K, B, T, accum, windows = 24, 160, 3, 6, 2, 2
    student, teacher = DinoStudentTeacher(_BB(), out_dim=K), DinoStudentTeacher(_BB(), out_dim=K)
    with torch.no_grad():
        for p in teacher.parameters():
            p.add_(torch.randn(p.shape) * 0.05)
            p.requires_grad_(False)
    dl = DINOLoss(K, 0.9)
    opt = torch.optim.AdamW(student.head.parameters(), lr=1e-2, weight_decay=0.04)
    g = torch.Generator().manual_seed(12)
    out = {"accum": accum, "windows": windows, "gram_weight": 0.5, "ema": 0.99, "lr": 1e-2, "weight_decay": 0.04,
           "student_temp": 0.1, "teacher_temp": 0.04, "momentum": 0.9}
    for k, v in student.head.state_dict().items():
        out[f"s0_{k.replace('.', '_')}"] = _np(v)
    for k, v in teacher.head.state_dict().items():
        out[f"t0_{k.replace('.', '_')}"] = _np(v)
    step = 0
    for w in range(windows):
        opt.zero_grad(set_to_none=True)
        for _ in range(accum):
            s_feats = torch.randn(2 * B, T, D, generator=g)
            t_feats = torch.randn(2 * B, T, D, generator=g)
            out[f"s_feats_{step}"], out[f"t_feats_{step}"] = _np(s_feats), _np(t_feats)
            student_out = student.head(s_feats[:, 0])
            with torch.no_grad():
                teacher_out = teacher.head(t_feats[:, 0])
            loss_dino = dl(student_out, teacher_out, 0.1, 0.04)
            loss_gram = compute_gram_anchoring_loss(s_feats, t_feats)
            ((loss_dino + 0.5 * loss_gram) / accum).backward()
            out[f"loss_dino_{step}"], out[f"loss_gram_{step}"] = _np(loss_dino), _np(loss_gram)
            out[f"center_{step}"] = _np(dl.center)
            step += 1
        for n, p in student.head.named_parameters():
            out[f"grad_w{w}_{n.replace('.', '_')}"] = _np(p.grad)
        opt.step()
        with torch.no_grad():
            for p_s, p_t in zip(student.parameters(), teacher.parameters()):
                p_t.data.mul_(0.99).add_(p_s.data, alpha=1.0 - 0.99)
        for k, v in student.head.state_dict().items():
            out[f"s_w{w}_{k.replace('.', '_')}"] = _np(v)
        for k, v in teacher.head.state_dict().items():
            out[f"t_w{w}_{k.replace('.', '_')}"] = _np(v)
    np.savez(os.path.join(OUT, "window_sequence.npz"), **out)
    print("  window_sequence.npz:", os.path.getsize(os.path.join(OUT, "window_sequence.npz")), "bytes")


def koleo_golden():
    """KoLeoLoss of the reference (scripts/phase5_big_run.py:742-773) with its autograd gradient on
    spread-out, clustered (near-duplicate rows: the cancellation-prone case) and wide-K inputs."""
    sys.path[:0] = [REF, os.path.join(REF, "scripts")]
    from phase5_big_run import KoLeoLoss
    out = {}
    for name, (r, k, seed, cluster) in {"small": (8, 128, 301, 0.0), "mid": (64, 1000, 302, 0.0),
                                        "clustered": (48, 512, 303, 0.05), "wide": (16, 8192, 304, 0.0)}.items():
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(r, k, generator=g) * 3.0
        if cluster:
            x = x[: r // 4].repeat(4, 1) + cluster * torch.randn(r, k, generator=g)
        xg = x.clone().requires_grad_(True)
        loss = KoLeoLoss()(xg)
        loss.backward()
        out[f"{name}_x"], out[f"{name}_loss"], out[f"{name}_grad"] = _np(x), _np(loss), _np(xg.grad)
    np.savez(os.path.join(OUT, "koleo.npz"), **out)
    print("koleo.npz written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "koleo":
        koleo_golden()      # adds one file, leaves the other vectors untouched
    elif len(sys.argv) > 1 and sys.argv[1] == "windows":
        os.makedirs(OUT, exist_ok=True)
        window_sequence_golden()
    else:
        main()
        koleo_golden()
        window_sequence_golden()

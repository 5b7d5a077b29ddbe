"""Pins oracle/losshead_oracle.py against golden vectors produced by the REFERENCE itself
(oracle/gen_golden.py, run in the build container) and checks the degenerate-case identities
that tie the extensions E1-E3 back to the reference."""
import math

import numpy as np
import pytest
import torch

from oracle import losshead_oracle as O

T = torch.from_numpy


def close(a, b, rtol=1e-6, atol=1e-7):
    a = torch.as_tensor(a, dtype=torch.float32)
    b = torch.as_tensor(b, dtype=torch.float32)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a.flatten()[:4], b.flatten()[:4], (a - b).abs().max())


def test_uniform_logits_entropy_wall(golden):
    g = golden("dino_uniform.npz")
    k = int(g["out_dim"])
    loss = O.dino_loss_reference(torch.zeros(4, k), torch.zeros(4, k), torch.zeros(1, k), 0.1, 0.04)
    close(loss, g["loss"])
    assert abs(loss.item() - math.log(k)) < 1e-5


def test_seeded_dino_two_calls(golden):
    g = golden("dino_seeded.npz")
    s, t = T(g["student"]), T(g["teacher"])
    center = torch.zeros(1, s.shape[1])
    for tag in ("a", "b"):
        sg = s.clone().requires_grad_(True)
        loss = O.dino_loss_reference(sg, t, center, 0.1, 0.04)
        loss.backward()
        center = O.center_update(center, t, 0.9)
        close(loss, g[f"loss_{tag}"])
        close(sg.grad, g[f"grad_{tag}"], rtol=1e-5, atol=1e-8)
        close(center, g[f"center_{tag}"])
    # SURVEY 8c values (transcription check of the recipe itself)
    assert abs(float(g["loss_a"]) - 28.66371536) < 1e-5
    assert abs(float(g["loss_b"]) - 28.93894005) < 1e-5


def test_multicrop_reduces_to_reference(golden):
    g = golden("dino_seeded.npz")
    s, t = T(g["student"]), T(g["teacher"])
    c = torch.zeros(1, s.shape[1])
    l_ref = O.dino_loss_reference(s, t, c, 0.1, 0.04)
    l_mc = O.multicrop_dino_loss(s, t, c, 0.1, 0.04, n_global=2, n_local=0)
    close(l_mc, l_ref, rtol=1e-6)
    close(l_mc, g["loss_a"], rtol=1e-6)


def test_analytic_gradient_multicrop():
    """dL/ds[v,b] = (1/(tau_s n_terms B)) * sum_{iq != v} (softmax(s/tau_s) - q[iq,b])."""
    g = torch.Generator().manual_seed(3)
    B, Vg, Vl, K = 3, 2, 3, 50
    V = Vg + Vl
    s = torch.randn(B * V, K, generator=g, requires_grad=True)
    t = torch.randn(B * Vg, K, generator=g)
    c = torch.randn(1, K, generator=g) * 0.1
    loss = O.multicrop_dino_loss(s, t, c, 0.1, 0.04, Vg, Vl)
    loss.backward()
    q = torch.softmax((t - c) / 0.04, -1).view(Vg, B, K)
    p = torch.softmax(s.detach() / 0.1, -1).view(V, B, K)
    n_terms = Vg * V - Vg
    grad = torch.zeros(V, B, K)
    for v in range(V):
        for iq in range(Vg):
            if iq != v:
                grad[v] += (p[v] - q[iq])
    grad = grad / (0.1 * n_terms * B)
    close(s.grad, grad.view(B * V, K), rtol=1e-4, atol=1e-7)


def test_dino_shapes(golden):
    g = golden("dino_shapes.npz")
    for name in ("k1000", "k4099", "k65536"):
        rows, k = [int(v) for v in g[f"{name}_shape"]]
        gen = torch.Generator().manual_seed(int(g[f"{name}_seed"]))
        scale = float(g[f"{name}_scale"])
        s = torch.randn(rows, k, generator=gen) * scale
        t = torch.randn(rows, k, generator=gen) * scale
        c0 = torch.randn(1, k, generator=gen) * 0.1
        if f"{name}_student" in g:
            close(s, g[f"{name}_student"], rtol=0, atol=0)
        sg = s.clone().requires_grad_(True)
        loss = O.dino_loss_reference(sg, t, c0, 0.1, 0.04)
        loss.backward()
        close(loss, g[f"{name}_loss"], rtol=2e-6)
        c1 = O.center_update(c0, t, float(g[f"{name}_mom"]))
        if f"{name}_grad" in g:
            close(sg.grad, g[f"{name}_grad"], rtol=1e-5, atol=1e-9)
            close(c1, g[f"{name}_center1"])
        else:
            close(sg.grad.norm(), g[f"{name}_grad_norm"], rtol=1e-5)
            close(sg.grad[:, :64], g[f"{name}_grad_head"], rtol=1e-5, atol=1e-9)
            close(c1[:, :64], g[f"{name}_center1"])
            close(c1.sum(), g[f"{name}_center1_sum"], rtol=1e-5, atol=1e-5)


def test_gram_and_koleo(golden):
    g = golden("dino_seeded.npz")
    sf, tf = T(g["gram_student"]), T(g["gram_teacher"])
    sfg = sf.clone().requires_grad_(True)
    loss = O.gram_anchoring_loss(sfg, tf)
    loss.backward()
    close(loss, g["gram_loss"])
    close(sfg.grad, g["gram_grad"], rtol=1e-5, atol=1e-9)
    assert sfg.grad[:, 0].abs().max().item() == 0.0  # CLS row gets no gradient
    close(O.gram_matrix(sf[:, 1:]), g["gram_matrix"])
    close(O.gram_anchoring_loss(tf, tf), g["gram_self"])
    close(O.koleo_loss(T(g["koleo_x"])), g["koleo"])


def test_head_and_ema(golden):
    g = golden("head_ema.npz")
    p = O.HeadParams(T(g["head_0_weight"]), T(g["head_0_bias"]), T(g["head_2_weight"]), T(g["head_2_bias"]))
    close(O.head_forward(T(g["cls"]), p), g["head_out"], rtol=1e-6, atol=1e-7)
    assert list(g["head_keys"]) == ["head.0.weight", "head.0.bias", "head.2.weight", "head.2.bias"]
    n = int(g["n_params"])
    ps = [T(g[f"ps_{i}"]).clone() for i in range(n)]
    pt = [T(g[f"pt0_{i}"]).clone() for i in range(n)]
    O.ema_update(pt, ps, float(g["ema"]))
    for i in range(n):
        assert torch.equal(pt[i], T(g[f"pt1_{i}"])), i  # bit exact: same op order as the reference


def _microstep_oracle(g, policy="fp32"):
    sp = O.HeadParams(*[T(g[f"s_head_{k}"]).clone().requires_grad_(True)
                        for k in ("0_weight", "0_bias", "2_weight", "2_bias")])
    tp = O.HeadParams(*[T(g[f"t_head_{k}"]) for k in ("0_weight", "0_bias", "2_weight", "2_bias")])
    orc = O.LossHeadOracle(sp, tp, out_dim=256, center_momentum=float(g["momentum"]),
                           n_global=2, n_local=0, policy=policy)
    orc.center = T(g["center0"]).clone()
    sf = T(g["student_feats"]).clone().requires_grad_(True)
    tf = T(g["teacher_feats"])
    out = orc.step(sf[:, 0], tf[:, 0], float(g["student_temp"]), float(g["teacher_temp"]),
                   student_tok=sf, teacher_tok=tf, accum=int(g["accum"]))
    return orc, sp, sf, out


def test_whole_microstep_matches_reference_loop(golden):
    g = golden("microstep.npz")
    orc, sp, sf, out = _microstep_oracle(g)
    close(out["loss_dino"], g["loss_dino"], rtol=2e-6)
    close(out["loss_gram"], g["loss_gram"], rtol=2e-6)
    close(orc.center, g["center1"])
    close(sf.grad, g["d_student_feats"], rtol=1e-4, atol=1e-8)
    for k, t in zip(("0_weight", "0_bias", "2_weight", "2_bias"), sp.tensors()):
        close(t.grad, g[f"g_head_{k}"], rtol=1e-4, atol=1e-8)


def test_two_optimizer_windows_match_reference_loop(golden):
    """a9 over time (scripts/phase5_big_run.py:1738-1802): two windows of accum = 2 micro-steps run by the REFERENCE
    (its DINOLoss, Gram loss, head, torch.optim.AdamW, EMA loop) against the oracle driven the same way: losses and
    the centre after every micro-step, the accumulated head gradients at each window end, the student head after
    AdamW and the teacher head after the EMA."""
    g = golden("window_sequence.npz")
    keys = ("0_weight", "0_bias", "2_weight", "2_bias")
    sp = O.HeadParams(*[T(g[f"s0_{k}"]).clone().requires_grad_(True) for k in keys])
    tp = O.HeadParams(*[T(g[f"t0_{k}"]).clone() for k in keys])
    K = sp.w2.shape[0]
    accum, windows = int(g["accum"]), int(g["windows"])
    orc = O.LossHeadOracle(sp, tp, out_dim=K, center_momentum=float(g["momentum"]), n_global=2, n_local=0,
                           gram_weight=float(g["gram_weight"]), policy="fp32")
    opt = torch.optim.AdamW(sp.tensors(), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    step = 0
    for w in range(windows):
        opt.zero_grad(set_to_none=True)
        for _ in range(accum):
            sf, tf = T(g[f"s_feats_{step}"]).clone().requires_grad_(True), T(g[f"t_feats_{step}"])
            out = orc.step(sf[:, 0], tf[:, 0], float(g["student_temp"]), float(g["teacher_temp"]),
                           student_tok=sf, teacher_tok=tf, accum=accum)
            close(out["loss_dino"], g[f"loss_dino_{step}"], rtol=2e-6)
            close(out["loss_gram"], g[f"loss_gram_{step}"], rtol=2e-6)
            close(orc.center, g[f"center_{step}"])
            step += 1
        for k, p in zip(keys, sp.tensors()):
            close(p.grad, g[f"grad_w{w}_{k}"], rtol=1e-4, atol=1e-8)
        opt.step()
        with torch.no_grad():
            O.ema_update(tp.tensors(), [p.detach() for p in sp.tensors()], float(g["ema"]))
        for k, p, q in zip(keys, sp.tensors(), tp.tensors()):
            close(p, g[f"s_w{w}_{k}"], rtol=1e-5, atol=1e-7)
            close(q, g[f"t_w{w}_{k}"], rtol=1e-6, atol=1e-7)


def test_bf16_policy_is_close_to_fp32(golden):
    """The bf16-operand policy (what the CUDA tensor-core path computes) stays within the
    north_star tolerance (rtol 1e-3 on the loss) of the fp32 reference on identical inputs."""
    g = golden("microstep.npz")
    _, _, _, out = _microstep_oracle(g, policy="bf16")
    assert abs(float(out["loss_dino"]) - float(g["loss_dino"])) / float(g["loss_dino"]) < 1e-3
    assert abs(float(out["loss_gram"]) - float(g["loss_gram"])) / float(g["loss_gram"]) < 5e-3


def test_sinkhorn_properties_and_dp_equivalence():
    """E2 is unpinned by the reference; check the defining properties: rows sum to 1, prototype
    mass is (nearly) uniform, and the result is invariant to a common shift of the logits."""
    g = torch.Generator().manual_seed(11)
    t = torch.randn(16, 64, generator=g)
    q = O.sinkhorn_knopp(t, 0.04, 3)
    close(q.sum(-1), torch.ones(16), rtol=1e-5)
    col = q.sum(0)
    col_sm = torch.softmax(t / 0.04, -1).sum(0)
    assert (col.max() / col.min()).item() < 0.1 * (col_sm.max() / col_sm.min()).item()
    close(O.sinkhorn_knopp(t + 3.0, 0.04, 3), q, rtol=1e-4, atol=1e-7)
    assert torch.isfinite(O.sinkhorn_knopp(t * 50.0, 0.04, 3)).all()


def test_sinkhorn_log_domain_oracle_equals_the_published_probability_domain_form():
    """E2 has no reference implementation to pin against; what can be pinned is the ALGORITHM: the published
    DINOv2 / SwAV iteration written literally in the probability domain in float64 numpy (Q = exp(t/tau)^T; Q /= sum Q;
    3 x { Q /= rowsum; Q /= K; Q /= colsum; Q /= B }; Q *= B) against the oracle's log-domain evaluation, on one
    process and with the batch split over two "ranks" whose per-prototype sums are added (the all-reduce)."""
    g = torch.Generator().manual_seed(21)
    t = (torch.randn(24, 96, generator=g) * 0.3).double().numpy()   # small enough that exp(t / 0.04) stays in range
    tau, n_it = 0.04, 3

    def published(tt, world_split=None):
        Q = np.exp(tt / tau).T                       # (K, B)
        K, B = Q.shape
        Q /= Q.sum()
        for _ in range(n_it):
            if world_split is None:
                rows = Q.sum(axis=1, keepdims=True)
            else:                                    # per-rank partial sums, then the all-reduce
                rows = sum(Q[:, sl].sum(axis=1, keepdims=True) for sl in world_split)
            Q /= rows
            Q /= K
            Q /= Q.sum(axis=0, keepdims=True)
            Q /= B
        Q *= B
        return Q.T

    ref = published(t.copy())
    got = O.sinkhorn_knopp(torch.from_numpy(t), tau, n_it).double().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * ref.max(), np.abs(got - ref).max()
    ref2 = published(t.copy(), world_split=[slice(0, 10), slice(10, 24)])
    assert np.abs(ref2 - ref).max() <= 1e-12
    np.testing.assert_allclose(ref.sum(1), 1.0, rtol=1e-12)


def test_multicrop_oracle_equals_the_published_dino_v1_loop():
    """E1: the public DINO-v1 loss loop written literally (teacher chunks x student chunks, skip v == iq, mean over the
    batch, divide by the number of terms; centre subtracted before the teacher softmax) in float64 against the oracle."""
    g = torch.Generator().manual_seed(22)
    B, Vg, Vl, K = 5, 2, 6, 48
    s = torch.randn((Vg + Vl) * B, K, generator=g).double()
    t = torch.randn(Vg * B, K, generator=g).double()
    c = (torch.randn(1, K, generator=g) * 0.1).double()
    student = (s / 0.1).chunk(Vg + Vl)
    teacher = torch.softmax((t - c) / 0.04, dim=-1).chunk(Vg)
    total, n_terms = 0.0, 0
    for iq, q in enumerate(teacher):
        for v in range(len(student)):
            if v == iq:
                continue
            total = total + torch.sum(-q * torch.log_softmax(student[v], dim=-1), dim=-1).mean()
            n_terms += 1
    total = total / n_terms
    got = O.multicrop_dino_loss(s.float(), t.float(), c.float(), 0.1, 0.04, Vg, Vl)
    assert n_terms == Vg * (Vg + Vl) - Vg
    close(got, total.float(), rtol=2e-6)


def test_ibot_oracle_equals_the_published_forward_masked_form():
    """E3: DINOv2's iBOTPatchLoss.forward_masked written literally in float64 (per-token CE, weighted by
    1 / n_masked(image), summed, divided by the number of images) against the oracle, with ragged mask counts."""
    g = torch.Generator().manual_seed(23)
    counts = [3, 7, 1, 5]                                    # masked tokens per image (global crop)
    Mm, K = sum(counts), 40
    s = torch.randn(Mm, K, generator=g).double()
    t = torch.randn(Mm, K, generator=g).double()
    c = (torch.randn(1, K, generator=g) * 0.1).double()
    w = torch.cat([torch.full((n,), 1.0 / n) for n in counts]).double()
    tprob = torch.softmax((t - c) / 0.04, dim=-1)
    loss = torch.sum(tprob * torch.log_softmax(s / 0.1, dim=-1), dim=-1) * w
    ref = -loss.sum() / len(counts)
    got = O.ibot_patch_loss(s.float(), t.float(), c.float(), 0.1, 0.04, w.float(), n_images=len(counts))
    close(got, ref.float(), rtol=2e-6)


def test_ibot_reduces_to_plain_ce():
    g = torch.Generator().manual_seed(12)
    s = torch.randn(10, 40, generator=g)
    t = torch.randn(10, 40, generator=g)
    c = torch.zeros(1, 40)
    w = torch.full((10,), 1.0 / 5)  # 2 images x 5 masked tokens
    l = O.ibot_patch_loss(s, t, c, 0.1, 0.04, w, n_images=2)
    q = torch.softmax(t / 0.04, -1)
    ref = -(q * torch.log_softmax(s / 0.1, -1)).sum(-1).mean()
    close(l, ref, rtol=1e-6)


def test_entropy_diagnostics_float64():
    g = torch.Generator().manual_seed(13)
    s = torch.randn(6, 100, generator=g)
    t = torch.randn(6, 100, generator=g)
    c = torch.randn(1, 100, generator=g) * 0.1
    te, se = O.entropy_diagnostics(s, t, c, 0.1, 0.04)
    pt = torch.softmax((t.double() - c.double()) / 0.04, -1)
    ps = torch.softmax(s.double() / 0.1, -1)
    close(te, -(pt * torch.log(pt.clamp_min(1e-300))).sum(-1).mean().float(), rtol=1e-5)
    close(se, -(ps * torch.log(ps.clamp_min(1e-300))).sum(-1).mean().float(), rtol=1e-5)


@pytest.mark.parametrize("name", ["small", "mid", "clustered", "wide"])
def test_koleo_loss_and_gradient(golden, name):
    """a11 KoLeo (scripts/phase5_big_run.py:742-773): oracle restatement against the reference's own
    loss and autograd gradient, including near-duplicate rows (small nearest-neighbour distances)."""
    g = golden("koleo.npz")
    x = T(g[f"{name}_x"]).clone().requires_grad_(True)
    loss = O.koleo_loss(x)
    loss.backward()
    close(loss, g[f"{name}_loss"])
    close(x.grad, g[f"{name}_grad"], rtol=1e-5, atol=1e-9)

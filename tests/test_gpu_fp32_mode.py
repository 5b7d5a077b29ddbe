"""GPU tests of the fp32-faithful contraction mode (the reference WITHOUT --amp, its default: nn.Linear and
torch.bmm in true fp32, scripts/phase5_big_run.py:1322).  Outside torch.autocast the drop-in modules evaluate every
contraction as three bf16 tensor-core GEMMs on hi/lo splits of the fp32 operands (~16 mantissa bits per product).
Against the fp32 golden vectors produced by the reference's own classes: losses and logits 2e-5, gradients 1e-4
relative L2 (the bf16 mode needs 1e-3 / 1.2e-2 on the same vectors)."""
import numpy as np
import pytest
import torch

from oracle import losshead_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture()
def dx():
    from dinox_b200 import losshead, _ext
    assert _ext.lib().dinox_device_check() == 0, "needs a B200"
    prev = losshead.set_contraction_precision("auto")     # the shipped default
    yield losshead
    losshead.set_contraction_precision(prev)


def test_auto_mode_follows_autocast(dx, golden):
    g = golden("head_ema.npz")
    head = dx.ProjectionHead(32, 96).to(DEV)
    head.load_state_dict({k: T(g["head_" + k.replace(".", "_")]) for k in head.state_dict().keys()})
    x = T(g["cls"]).to(DEV)
    out32 = head(x)                                          # no autocast: fp32-faithful
    assert out32.dtype == torch.float32 and rel(out32, T(g["head_out"])) < 2e-5
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):   # the reference's --amp: bf16 operands, bf16 logits
        outb = head(x)
    assert outb.dtype == torch.bfloat16 and 1e-4 < rel(outb, T(g["head_out"])) < 1e-2
    with dx.contraction_precision("bf16"):
        assert rel(head(x), T(g["head_out"])) > 1e-4


def test_head_fp32_mode_forward_backward_vs_fp32_oracle(dx):
    from dinox_b200 import synth
    gen = torch.Generator().manual_seed(10)
    rows, D, K = 200, 384, 4099                              # ragged K, several M tiles
    sd = synth.head_weights(D, K, gen)
    x = torch.randn(rows, D, generator=gen)
    dz = torch.randn(rows, K, generator=gen) / K
    p = O.HeadParams(*[sd[k].clone().requires_grad_(True) for k in ("0.weight", "0.bias", "2.weight", "2.bias")])
    xo = x.clone().requires_grad_(True)
    z = O.head_forward(xo, p, policy="fp32")
    z.backward(dz)
    head = dx.ProjectionHead(D, K).to(DEV)
    head.load_state_dict(sd)
    xd = x.to(DEV).requires_grad_(True)
    zd = head(xd)
    zd.backward(dz.to(DEV))
    assert rel(zd, z) < 2e-5
    assert rel(xd.grad, xo.grad) < 1e-4
    for (n, q), r in zip(head.named_parameters(), p.tensors()):
        assert rel(q.grad, r.grad) < 1e-4, n


def test_gram_fp32_mode_golden(dx, golden):
    g = golden("dino_seeded.npz")
    sf = T(g["gram_student"]).to(DEV).requires_grad_(True)
    tf = T(g["gram_teacher"]).to(DEV)
    loss = dx.compute_gram_anchoring_loss(sf, tf)
    loss.backward()
    ref = float(g["gram_loss"])
    assert abs(loss.item() - ref) <= 2e-5 * ref, (loss.item(), ref)
    assert rel(sf.grad, T(g["gram_grad"])) < 1e-4
    assert sf.grad[:, 0].abs().max().item() == 0.0
    gm = dx.compute_gram_matrix(T(g["gram_student"])[:, 1:].contiguous().to(DEV))
    assert rel(gm, T(g["gram_matrix"])) < 2e-5


@pytest.mark.parametrize("shape", [(3, 201, 384), (2, 261, 64)])
def test_gram_fp32_mode_vs_oracle_shapes(dx, shape):
    gen = torch.Generator().manual_seed(9)
    sf = torch.randn(*shape, generator=gen)
    tf = sf + 0.3 * torch.randn(*shape, generator=gen)
    so = sf.clone().requires_grad_(True)
    ref = O.gram_anchoring_loss(so, tf, policy="fp32")
    (ref * 3.0).backward()
    sd = sf.to(DEV).requires_grad_(True)
    loss = dx.compute_gram_anchoring_loss(sd, tf.to(DEV))
    (loss * 3.0).backward()
    assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item())
    assert rel(sd.grad, so.grad) < 1e-4
    # compute_gram_matrix with its own backward
    x = sf[:1, 1:].contiguous()
    xo = x.clone().requires_grad_(True)
    w = torch.randn(1, shape[1] - 1, shape[1] - 1, generator=gen)
    (O.gram_matrix(xo, "fp32") * w).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    (dx.compute_gram_matrix(xd) * w.to(DEV)).sum().backward()
    assert rel(xd.grad, xo.grad) < 1e-4


def test_microstep_golden_fp32_mode(dx, golden):
    """The reference call sequence (scripts/phase5_big_run.py:1746-1772) without --amp, against the golden produced
    by the reference's own classes in fp32."""
    g = golden("microstep.npz")
    K, D = 256, 32
    s_head, t_head = dx.ProjectionHead(D, K).to(DEV), dx.ProjectionHead(D, K).to(DEV)
    s_head.load_state_dict({k: T(g["s_head_" + k.replace(".", "_")]) for k in s_head.state_dict()})
    t_head.load_state_dict({k: T(g["t_head_" + k.replace(".", "_")]) for k in t_head.state_dict()})
    dl = dx.DINOLoss(K, float(g["momentum"])).to(DEV)
    dl.center.copy_(T(g["center0"]))
    sf = T(g["student_feats"]).to(DEV).requires_grad_(True)
    tf = T(g["teacher_feats"]).to(DEV)
    accum = int(g["accum"])
    student_out = s_head(sf[:, 0])
    teacher_out = t_head(tf[:, 0])
    loss_dino = dl(student_out, teacher_out, float(g["student_temp"]), float(g["teacher_temp"]))
    loss_gram = dx.compute_gram_anchoring_loss(sf, tf)
    ((loss_dino + 1.0 * loss_gram) / accum).backward()
    assert abs(loss_dino.item() - float(g["loss_dino"])) <= 2e-5 * float(g["loss_dino"])
    assert abs(loss_gram.item() - float(g["loss_gram"])) <= 2e-5 * float(g["loss_gram"])
    assert rel(dl.center, T(g["center1"])) < 2e-5
    assert rel(sf.grad, T(g["d_student_feats"])) < 1e-4
    for k in ("0_weight", "0_bias", "2_weight", "2_bias"):
        q = dict(s_head.named_parameters())[k.replace("_", ".")]
        assert rel(q.grad, T(g[f"g_head_{k}"])) < 1e-4, k

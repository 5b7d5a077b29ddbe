"""Regression tests for the advisor findings of round 1:
  * bf16 operand copies of the head weights must follow parameter updates that do not bump tensor version
    counters: the reference's own EMA loop (`p_t.data.mul_(m).add_(p_s.data, alpha=1-m)`,
    scripts/phase5_big_run.py:1800-1802) and FusedAdamW's raw-pointer kernel,
  * padding entries (entry count not a multiple of 128: the reference default of 2 views with B = 32) next to large
    biases / centres must not produce 0 * inf,
  * ShardedFusedAdamW speaks torch.optim's param_groups / state_dict / load_state_dict."""
import math

import pytest
import torch

from oracle import losshead_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture()
def dx():
    from dinox_b200 import losshead, _ext
    assert _ext.lib().dinox_device_check() == 0, "needs a B200"
    prev = losshead.set_weight_cache("always")
    yield losshead
    losshead.set_weight_cache(prev)


def _heads(dx, D, K, seed):
    from dinox_b200 import synth
    gen = torch.Generator().manual_seed(seed)
    s, t = dx.ProjectionHead(D, K).to(DEV), dx.ProjectionHead(D, K).to(DEV)
    s.load_state_dict(synth.head_weights(D, K, gen))
    t.load_state_dict(synth.head_weights(D, K, gen))
    for p in t.parameters():
        p.requires_grad_(False)
    return s, t, gen


@pytest.mark.parametrize("mode", ["always", "tracked"])
def test_reference_ema_loop_and_fused_adamw_refresh_the_operand_copies(dx, mode):
    from dinox_b200 import FusedAdamW
    dx.set_weight_cache(mode)
    B, D, K = 8, 64, 512
    s_head, t_head, gen = _heads(dx, D, K, 3)
    xs, xt = torch.randn(2 * B, D, generator=gen).to(DEV), torch.randn(2 * B, D, generator=gen).to(DEV)

    def loss_now(sh, th):
        dl = dx.DINOLoss(K, 0.9).to(DEV)
        return dx.fused_head_dino_loss(xs, xt, sh, th, dl, 0.1, 0.04)["loss"].item()

    l0 = loss_now(s_head, t_head)
    # the reference's EMA loop with m = 0: teacher <- student, through .data (no version bump)
    with torch.no_grad():
        for p_s, p_t in zip(s_head.parameters(), t_head.parameters()):
            p_t.data.mul_(0.0).add_(p_s.data, alpha=1.0)
    if mode == "tracked":
        dx.invalidate_weight_cache()          # the documented duty of a loop that bypasses the dinox entry points
    l1 = loss_now(s_head, t_head)
    fresh_s, fresh_t, _ = _heads(dx, D, K, 3)
    fresh_t.load_state_dict(s_head.state_dict())
    l1_ref = loss_now(fresh_s, fresh_t)
    assert l1 != l0 and abs(l1 - l1_ref) <= 1e-6 * abs(l1_ref), (l0, l1, l1_ref)
    # FusedAdamW writes the student through raw pointers: the next forward must see the new weights in BOTH modes
    opt = FusedAdamW(s_head.parameters(), lr=0.05, weight_decay=0.0)
    x = xs.clone().requires_grad_(True)
    dl = dx.DINOLoss(K, 0.9).to(DEV)
    dx.fused_head_dino_loss(x, xt, s_head, t_head, dl, 0.1, 0.04)["loss"].backward()
    opt.step()
    l2 = loss_now(s_head, t_head)
    fresh_s.load_state_dict(s_head.state_dict())
    l2_ref = loss_now(fresh_s, fresh_t)
    assert l2 != l1 and abs(l2 - l2_ref) <= 1e-6 * abs(l2_ref), (l1, l2, l2_ref)


@pytest.mark.parametrize("pass2", ["readback", "recompute"])
def test_padding_entries_with_large_offsets_stay_finite(dx, pass2, monkeypatch):
    """2 views, B = 32: 64 CLS entries padded to 128.  A student bias of +12 on some prototypes and a centre that puts
    (b2t - c)/tau_t far below -128 log2 units made 2^x overflow on the padding entries (0 * inf = NaN)."""
    monkeypatch.setenv("DINOX_PASS2", pass2)
    B, D, K = 32, 64, 1024
    s_head, t_head, gen = _heads(dx, D, K, 5)
    with torch.no_grad():
        s_head[2].bias[:17] += 12.0
        t_head[2].bias[100:140] -= 9.0
    c0 = torch.zeros(1, K)
    c0[0, 200:260] = 6.0
    feats = dict(student_cls=torch.randn(2 * B, D, generator=gen), teacher_cls=torch.randn(2 * B, D, generator=gen))
    sp = O.HeadParams(*[p.detach().cpu().clone().requires_grad_(True) for p in s_head.parameters()])
    tp = O.HeadParams(*[p.detach().cpu().clone() for p in t_head.parameters()])
    orc = O.LossHeadOracle(sp, tp, K, center_momentum=0.9, policy="bf16")
    orc.center = c0.clone()
    xo = feats["student_cls"].clone().requires_grad_(True)
    ref = orc.step(xo, feats["teacher_cls"], 0.1, 0.04)
    dl = dx.DINOLoss(K, 0.9).to(DEV)
    dl.center.copy_(c0)
    x = feats["student_cls"].to(DEV).requires_grad_(True)
    out = dx.fused_head_dino_loss(x, feats["teacher_cls"].to(DEV), s_head, t_head, dl, 0.1, 0.04)
    out["loss"].backward()
    torch.cuda.synchronize()
    assert math.isfinite(out["loss"].item())
    for p in s_head.parameters():
        assert torch.isfinite(p.grad).all()
    assert abs(out["loss"].item() - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())
    assert rel(x.grad, xo.grad) < 4e-3
    assert rel(s_head[2].weight.grad, sp.w2.grad) < 4e-3 and rel(s_head[2].bias.grad, sp.b2.grad) < 4e-3


def test_sharded_adamw_optimizer_protocol(dx):
    """param_groups (lr schedule writes), state_dict / load_state_dict round trip in torch.optim.AdamW's layout."""
    from dinox_b200 import ShardedFusedAdamW
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(256, 64, device=DEV)), torch.nn.Parameter(torch.randn(64, device=DEV))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = ShardedFusedAdamW(ps, lr=1e-2, weight_decay=0.01)
    topt = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.01)
    for step in range(3):
        lr = 1e-2 * (step + 1)
        opt.param_groups[0]["lr"] = lr                      # scripts/phase5_big_run.py:1699
        topt.param_groups[0]["lr"] = lr
        for p, r in zip(ps, ref):
            g = torch.randn_like(p)
            p.grad, r.grad = g.clone(), g.clone()
        opt.step()
        topt.step()
    for p, r in zip(ps, ref):
        assert rel(p, r) < 1e-6
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["lr"] == pytest.approx(3e-2)
    topt2 = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0)
    topt2.load_state_dict(sd)                               # torch reads what we wrote
    opt2 = ShardedFusedAdamW([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0)
    opt2.load_state_dict(topt.state_dict())                 # and we read what torch wrote
    assert opt2.step_count == 3 and opt2.param_groups[0]["lr"] == pytest.approx(3e-2)
    for q, r in zip(opt2.params, ref):
        g = torch.randn_like(q)
        q.grad, r.grad = g.clone(), g.clone()
    opt2.step()
    topt.step()
    for q, r in zip(opt2.params, ref):
        assert rel(q, r) < 1e-6


def test_head_gradients_are_visible_to_autograd(dx):
    """Default fused path: dW1/db1/dW2/db2 are autograd gradients of the head parameters (torch.autograd.grad and
    parameter hooks see them); grads_in_place=True writes the same values straight into .grad instead."""
    B, D, K = 8, 64, 1024
    s_head, t_head, gen = _heads(dx, D, K, 7)
    xs, xt = torch.randn(2 * B, D, generator=gen).to(DEV), torch.randn(2 * B, D, generator=gen).to(DEV)
    seen = []
    hook = s_head[2].weight.register_hook(lambda g: seen.append(g.detach().clone()))
    x = xs.clone().requires_grad_(True)
    dl = dx.DINOLoss(K, 0.9).to(DEV)
    loss = dx.fused_head_dino_loss(x, xt, s_head, t_head, dl, 0.1, 0.04, update_center=False)["loss"]
    grads = torch.autograd.grad(loss, [x] + list(s_head.parameters()))
    hook.remove()
    assert len(seen) == 1 and torch.equal(seen[0], grads[3])
    assert all(p.grad is None for p in s_head.parameters())        # autograd.grad does not touch .grad
    # the nn.Module form and the in-place fast path give the same numbers
    mod = dx.FusedLossHead(s_head, t_head, dx.DINOLoss(K, 0.9).to(DEV)).to(DEV)
    x2 = xs.clone().requires_grad_(True)
    mod(x2, xt, 0.1, 0.04).backward()
    for p, g in zip(s_head.parameters(), grads[1:]):
        assert torch.equal(p.grad, g)
        p.grad = None
    x3 = xs.clone().requires_grad_(True)
    dx.fused_head_dino_loss(x3, xt, s_head, t_head, dx.DINOLoss(K, 0.9).to(DEV), 0.1, 0.04, update_center=False,
                            grads_in_place=True)["loss"].backward()
    for p, g in zip(s_head.parameters(), grads[1:]):
        assert rel(p.grad, g) < 1e-6
    assert torch.equal(x3.grad, grads[0]) and torch.equal(x2.grad, grads[0])

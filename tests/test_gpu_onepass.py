"""GPU parity of the ONE-pass cross-entropy forward on materialised logits (`dinox_ce_fwd_onepass`): against the
CPU oracle (scripts/phase5_big_run.py:703-717 restated in oracle/losshead_oracle.py), against the three-pass
kernels it replaces (rows_lse x2 + ce_fwd), and through the drop-in `DINOLoss` in both forward modes.
fp32 row math on both sides: rtol 1e-5 on the loss, 1e-5 relative L2 on LSE vectors and gradients."""
import math

import pytest
import torch

from oracle import losshead_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from dinox_b200 import _ext, ops
    assert _ext.lib().dinox_device_check() == 0, "needs a B200"
    return ops


CASES = [
    # B, V, Vg, K, dtype            what
    (4, 2, 2, 128, torch.float32),      # the reference's two views, one K-split
    (5, 6, 2, 1003, torch.float32),     # ragged K: scalar tail path
    (3, 10, 2, 4099, torch.float32),    # C2's crop layout, ragged K, several splits
    (8, 12, 4, 2048, torch.float32),    # the most views / global views the kernel takes
    (7, 3, 1, 8192, torch.float32),     # one global view
    (16, 10, 2, 65536, torch.float32),  # full K
    (6, 10, 2, 8192, torch.bfloat16),   # autocast logits
    (6, 4, 2, 8192, torch.float16),
]


@pytest.mark.parametrize("B,V,Vg,K,dtype", CASES)
def test_onepass_vs_oracle_and_three_passes(ops, B, V, Vg, K, dtype):
    gen = torch.Generator().manual_seed(100 * B + V + K)
    s = (torch.randn(V * B, K, generator=gen) * 1.5).to(dtype)
    t = (torch.randn(Vg * B, K, generator=gen) * 1.5).to(dtype)
    c = torch.randn(1, K, generator=gen) * 0.1
    ts, tt = 0.1, 0.04
    ref = O.multicrop_dino_loss(s.float(), t.float(), c, ts, tt, Vg, V - Vg)
    sd, td = s.to(DEV), t.to(DEV)
    colb = ops.axpb(c.reshape(-1).to(DEV), 1 / tt)
    norm = 1.0 / ((Vg * V - Vg) * B)
    loss, lse_s, rowb = ops.ce_fwd_onepass(sd, td, B, V, Vg, 1 / ts, 1 / tt, colb, None, norm, True)
    # the three-pass kernels on the same inputs
    rowb3 = ops.rows_lse(td, 1 / tt, colb)
    lse3 = ops.rows_lse(sd, 1 / ts)
    loss3 = ops.ce_fwd(sd, td, B, V, Vg, 1 / ts, 1 / tt, colb, rowb3, lse3, None, norm, True)
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), (loss.item(), ref.item())
    assert abs(loss.item() - loss3.item()) <= 1e-5 * abs(loss3.item()), (loss.item(), loss3.item())
    assert rel(lse_s, lse3) <= 1e-6 and rel(rowb, rowb3) <= 1e-6
    lse_ref = torch.logsumexp(s.float() / ts, dim=-1)
    rowb_ref = torch.logsumexp((t.float() - c) / tt, dim=-1)
    assert rel(lse_s, lse_ref) <= 1e-6 and rel(rowb, rowb_ref) <= 1e-6
    # the by-products feed the backward kernel: gradient against autograd of the oracle
    so = s.float().clone().requires_grad_(True)
    O.multicrop_dino_loss(so, t.float(), c, ts, tt, Vg, V - Vg).backward()
    grad = ops.ce_bwd(sd, td, B, V, Vg, 1 / ts, 1 / tt, colb, rowb, lse_s, None, norm, True, torch.ones((), device=DEV))
    assert rel(grad, so.grad) <= (1e-5 if dtype == torch.float32 else 8e-3)   # 16-bit gradients are rounded on store


def test_onepass_ibot_form_with_weights(ops):
    """V = Vg = 1, all pairs, per-row weights: the masked-patch term on materialised rows."""
    gen = torch.Generator().manual_seed(7)
    Mm, K, n_img = 300, 2048, 4
    s = torch.randn(Mm, K, generator=gen)
    t = torch.randn(Mm, K, generator=gen)
    c = torch.randn(1, K, generator=gen) * 0.1
    w = torch.rand(Mm, generator=gen)
    ref = O.ibot_patch_loss(s, t, c, 0.1, 0.04, w, n_img)
    colb = ops.axpb(c.reshape(-1).to(DEV), 25.0)
    loss, lse_s, rowb = ops.ce_fwd_onepass(s.to(DEV), t.to(DEV), Mm, 1, 1, 10.0, 25.0, colb, w.to(DEV), 1.0 / n_img, False)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert rel(lse_s, torch.logsumexp(s / 0.1, -1)) <= 1e-6


def test_onepass_uniform_logits_is_ln_k(ops):
    """docs/phase5_big_run.md:375-377: uniform outputs => loss = ln K."""
    K, B = 8192, 4
    z = torch.zeros(2 * B, K, device=DEV)
    loss, _, _ = ops.ce_fwd_onepass(z, z, B, 2, 2, 10.0, 25.0, None, None, 1.0 / (2 * B), True)
    assert abs(loss.item() - math.log(K)) <= 1e-5 * math.log(K)


def test_onepass_large_logits_and_strided_rows(ops):
    """Running-maximum rescale: logits whose exponentials overflow fp32 without it; rows with a leading dimension > K."""
    gen = torch.Generator().manual_seed(3)
    B, V, Vg, K = 4, 4, 2, 1536
    s_full = torch.randn(V * B, K + 64, generator=gen) * 40.0
    t_full = torch.randn(Vg * B, K + 32, generator=gen) * 20.0
    s, t = s_full[:, :K], t_full[:, :K]
    c = torch.zeros(1, K)
    ref = O.multicrop_dino_loss(s.double(), t.double(), c.double(), 0.1, 0.04, Vg, V - Vg)
    sd, td = s_full.to(DEV)[:, :K], t_full.to(DEV)[:, :K]
    loss, lse_s, rowb = ops.ce_fwd_onepass(sd, td, B, V, Vg, 10.0, 25.0, None, None, 1.0 / ((Vg * V - Vg) * B), True)
    assert math.isfinite(loss.item())
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())   # fp32 sums of terms ~1e3 against float64
    assert rel(lse_s, torch.logsumexp(s.double() / 0.1, -1)) <= 1e-6


def test_onepass_nonfinite_input_propagates(ops):
    """A NaN logit must reach the loss (the reference's detect_anomaly, scripts/phase5_big_run.py:1216-1218)."""
    B, K = 2, 1024
    s = torch.randn(2 * B, K, device=DEV)
    t = torch.randn(2 * B, K, device=DEV)
    s[1, 17] = float("nan")
    loss, _, _ = ops.ce_fwd_onepass(s, t, B, 2, 2, 10.0, 25.0, None, None, 1.0 / (2 * B), True)
    assert not math.isfinite(loss.item())


def test_onepass_rejects_too_many_views(ops):
    from dinox_b200 import _ext
    V = ops.ce_onepass_max_views() + 1
    s = torch.randn(V * 2, 256, device=DEV)
    t = torch.randn(4, 256, device=DEV)
    with pytest.raises(_ext.DinoxError):
        ops.ce_fwd_onepass(s, t, 2, V, 2, 10.0, 25.0, None, None, 1.0, True)


@pytest.mark.parametrize("V", [2, 10, 14])
def test_dinoloss_forward_modes_agree(V):
    """The drop-in DINOLoss in "onepass" and "passes" mode: same loss, gradient and centre; V = 14 exceeds the one-pass
    kernel's view count and silently takes the three-pass kernels in both modes."""
    from dinox_b200 import losshead
    gen = torch.Generator().manual_seed(11 + V)
    B, Vg, K = 6, 2, 4096
    s = torch.randn(V * B, K, generator=gen)
    t = torch.randn(Vg * B, K, generator=gen)
    c0 = torch.randn(1, K, generator=gen) * 0.05
    out = {}
    for mode in ("onepass", "passes"):
        prev = losshead.set_ce_forward(mode)
        try:
            crit = losshead.DINOLoss(K, 0.9, n_global=Vg, n_local=V - Vg).to(DEV)
            crit.center.copy_(c0.to(DEV))
            sd = s.to(DEV).requires_grad_(True)
            loss = crit(sd, t.to(DEV), 0.1, 0.04)
            loss.backward()
            out[mode] = (loss.detach().cpu(), sd.grad.cpu(), crit.center.cpu().clone())
        finally:
            losshead.set_ce_forward(prev)
    so = s.clone().requires_grad_(True)
    ref = O.multicrop_dino_loss(so, t, c0, 0.1, 0.04, Vg, V - Vg)
    ref.backward()
    for mode, (loss, grad, centre) in out.items():
        assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), mode
        assert rel(grad, so.grad) <= 1e-5, mode
    assert rel(out["onepass"][1], out["passes"][1]) <= 5e-6   # two fp32 summation orders of the same LSEs
    assert torch.equal(out["onepass"][2], out["passes"][2])   # the centre update is the same kernel pair


def test_onepass_register_form_in_a_child_process():
    """Aligned shapes take the persistent shared-memory staged kernel; DINOX_CE_STREAM=0 (read once per process) keeps
    them on the register form, which must agree with the oracle just the same."""
    import os
    import subprocess
    import sys
    code = r'''
import math, torch
from oracle import losshead_oracle as O
from dinox_b200 import ops
for (B, V, Vg, K, dt) in ((4, 2, 2, 128, torch.float32), (5, 10, 2, 16384, torch.float32), (6, 10, 2, 8192, torch.bfloat16)):
    gen = torch.Generator().manual_seed(K + V)
    s = (torch.randn(V * B, K, generator=gen) * 1.5).to(dt); t = (torch.randn(Vg * B, K, generator=gen) * 1.5).to(dt)
    c = torch.randn(1, K, generator=gen) * 0.1
    ref = O.multicrop_dino_loss(s.float(), t.float(), c, 0.1, 0.04, Vg, V - Vg)
    colb = ops.axpb(c.reshape(-1).cuda(), 25.0)
    loss, lse_s, rowb = ops.ce_fwd_onepass(s.cuda(), t.cuda(), B, V, Vg, 10.0, 25.0, colb, None, 1.0 / ((Vg * V - Vg) * B), True)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), (B, V, K, loss.item(), ref.item())
    l = torch.logsumexp(s.float() / 0.1, -1)
    assert ((lse_s.cpu() - l).norm() / l.norm()).item() <= 1e-6
print("REGISTER FORM OK")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DINOX_CE_STREAM="0", PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REGISTER FORM OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

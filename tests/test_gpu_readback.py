"""GPU tests of the read-back pass pair (dinox_head_teacher -> dinox_head_grad2) through the C ABI.

Reference: plain PyTorch fp32 on the same bf16-rounded operands (the contraction is exact in fp32 up to
summation order).  Tolerances: teacher probabilities are stored as fp16 relative to their granule (128 or 64 prototypes)
maximum (2^-12 relative on every value within 14 binades of it): probabilities 1e-3 relative L2,
row statistics 1e-5; G is bf16 (2^-9 per element): relative L2 4e-3; loss 1e-4; db2 sums the bf16 G: 4e-3.
Also the edge cases: ragged rows / prototypes (K not a multiple of 128 or 256), an alternative column
offset from a given M tile on, padding entries with weight 0 next to huge offsets (no 0 * inf), D > 384
(no resident operand) and a single M tile (no CTA pair)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG2E = 1.4426950408889634


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from dinox_b200 import ops, _ext
    assert _ext.lib().dinox_device_check() == 0, "needs a B200"
    return ops


def _teacher_case(rows, K, D, seed, alt_from=None, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    h = (torch.randn(rows, D, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(K, D, generator=g) * spread / math.sqrt(D)).to(torch.bfloat16)
    col = torch.randn(K, generator=g)
    col_alt = torch.randn(K, generator=g) if alt_from is not None else None
    return h, w, col, col_alt


def _teacher_ref(h, w, inv_tau, col, col_alt, alt_from):
    x = (h.double() @ w.double().t()) * (inv_tau * LOG2E) + col.double()[None, :]
    if col_alt is not None:
        x[alt_from:] = (h[alt_from:].double() @ w.double().t()) * (inv_tau * LOG2E) + col_alt.double()[None, :]
    lse2 = torch.logsumexp(x * math.log(2.0), dim=1) / math.log(2.0)
    q = torch.exp2(x - lse2[:, None])
    return x, lse2, q


@pytest.mark.parametrize("rows,K,D,alt_from,spread", [
    (300, 1000, 384, None, 1.0),      # ragged rows and prototypes, CTA pair with a half-empty super tile
    (128, 4099, 384, None, 4.0),      # a single M tile (no pair); a granule with 3 valid prototypes and an empty one
    (512, 2048, 384, 256, 8.0),       # column offsets switch at M tile 2; wide logit range (sharp rows)
    (384, 1536, 1024, 128, 2.0),      # D > 384: no resident operand, per-tile partials
    (256, 65536, 128, None, 2.0),     # full prototype count
])
def test_head_teacher_probabilities_and_statistics(ops, rows, K, D, alt_from, spread):
    h, w, col, col_alt = _teacher_case(rows, K, D, 11, alt_from, spread)
    inv_tau = 25.0
    x, lse2_ref, q_ref = _teacher_ref(h, w, inv_tau, col, col_alt, alt_from)
    qt, refs, lse2 = ops.head_teacher(h.to(DEV), w.to(DEV), inv_tau, col.to(DEV),
                                      None if col_alt is None else col_alt.to(DEV), alt_from or 0)
    torch.cuda.synchronize()
    gpt, tile = ops.teacher_granules_per_tile(), ops.teacher_tile_cols()
    gw = tile // gpt                                                  # prototypes per granule
    assert qt.shape == (rows, (K + tile - 1) // tile * tile) and refs.shape == (gpt * ((K + tile - 1) // tile), rows)
    assert torch.isfinite(qt.float()).all()
    assert (qt[:, K:] == 0).all(), "padding prototypes must be written as zeros"
    assert rel(lse2, lse2_ref) < 1e-5
    # granule maxima
    ng = (K + gw - 1) // gw
    gmax_ref = torch.stack([x[:, g * gw:min(K, (g + 1) * gw)].max(dim=1).values for g in range(ng)], 0)
    assert rel(refs[:ng], gmax_ref) < 3e-6   # fp32 accumulation order over D
    # reconstruct q = qt * 2^(ref - lse2)
    scale = torch.exp2(refs[:ng].double().cpu() - lse2.double().cpu()[None, :])          # (ng, rows)
    q = qt[:, :K].double().cpu() * scale.t().repeat_interleave(gw, dim=1)[:, :K]
    assert rel(q, q_ref) < 1e-3
    assert (q.sum(1) - 1).abs().max() < 2e-3
    # every value within 14 binades of its granule maximum carries fp16-normal precision
    big = q_ref > q_ref.max(dim=1, keepdim=True).values * 2.0 ** -12
    assert ((q - q_ref).abs()[big] / q_ref[big]).max() < 1.5e-3


def _grad2_case(E, K, D, Tr, seed, n_pad=0):
    g = torch.Generator().manual_seed(seed)
    hs = (torch.randn(E, D, generator=g) * 0.5).to(torch.bfloat16)
    ht = (torch.randn(Tr, D, generator=g) * 0.5).to(torch.bfloat16)
    ws = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16)
    wt = (torch.randn(K, D, generator=g) * 3.0 / math.sqrt(D)).to(torch.bfloat16)
    cs = torch.randn(K, generator=g)
    ct = torch.randn(K, generator=g)
    cw = torch.rand(E, generator=g) / E
    trow = torch.randint(0, Tr, (E,), generator=g)
    if n_pad:
        hs[-n_pad:] = 0
        cw[-n_pad:] = 0
    return hs, ht, ws, wt, cs, ct, cw, trow


@pytest.mark.parametrize("E,K,D,Tr,n_pad,alt", [
    (256, 1000, 384, 70, 0, 128),
    (384, 4099, 384, 130, 100, 256),     # padding entries (weight 0, huge offsets), ragged K
    (128, 2048, 384, 128, 0, 1 << 30),   # single M tile, everything in loss[0]
    (256, 1536, 1024, 64, 17, 0),        # D > 384, everything in loss[1]
    (640, 65536, 128, 200, 0, 384),
])
def test_head_grad2_against_torch(ops, E, K, D, Tr, n_pad, alt):
    hs, ht, ws, wt, cs, ct, cw, trow = _grad2_case(E, K, D, Tr, 5, n_pad)
    inv_ts, inv_tt = 10.0, 25.0
    # teacher side through the teacher kernel (tested above)
    qt, refs, rb2_t = ops.head_teacher(ht.to(DEV), wt.to(DEV), inv_tt, ct.to(DEV))
    # student statistics
    u = (hs.double() @ ws.double().t()) * (inv_ts * LOG2E) + cs.double()[None, :]
    lse2_s = torch.logsumexp(u * math.log(2.0), dim=1) / math.log(2.0)
    p = torch.exp2(u - lse2_s[:, None])
    _, _, q_t = _teacher_ref(ht, wt, inv_tt, ct, None, None)
    q = q_t[trow]
    G_ref = (cw.double() * inv_ts)[:, None] * (p - q)
    per_entry = -(cw.double()[:, None] * q * (u - lse2_s[:, None]) * math.log(2.0)).sum(1)
    loss_ref = torch.stack([per_entry[:min(alt, E)].sum(), per_entry[min(alt, E):].sum()])
    lse2_e = lse2_s.float().clone()
    rb2_e = rb2_t[trow.to(DEV)].clone()
    if n_pad:   # what the host passes for padding entries
        lse2_e[-n_pad:] = 1.0e30
        rb2_e[-n_pad:] = 1.0e30
    losses = torch.full((3,), float("nan"), device=DEV)
    G, db2p = ops.head_grad2(hs.to(DEV), ws.to(DEV), inv_ts, cs.to(DEV), lse2_e.to(DEV), cw.to(DEV), rb2_e,
                             trow.to(torch.int32).to(DEV), qt, refs, alt, losses)
    torch.cuda.synchronize()
    assert G.shape == (E, K) and torch.isfinite(G.float()).all() and torch.isfinite(losses).all()
    assert rel(G, G_ref) < 4e-3
    if n_pad:
        assert (G[-n_pad:] == 0).all()
    tot_ref = loss_ref.sum()
    assert losses[2].item() == (losses[0] + losses[1]).item()          # pass 2 writes the total itself
    assert abs(losses[2].item() - tot_ref.item()) <= 1e-4 * abs(tot_ref.item())
    for i in range(2):
        assert abs(losses[i].item() - loss_ref[i].item()) <= 1e-4 * abs(tot_ref.item()) + 1e-7
    db2 = db2p.double().sum(0).cpu()
    assert db2p.shape == ((E + 31) // 32, K)
    assert rel(db2, G_ref.sum(0)) < 4e-3
    # the bias gradient is the column sum of the bf16 G that the dW2 GEMM consumes
    assert rel(db2, G.double().sum(0).cpu()) < 1e-5


def test_head_grad2_matches_recompute_pass2(ops):
    """Same entries through the round-1 pass 2 (both logit tiles recomputed): G^T and the loss agree."""
    E, K, D, Tr = 512, 8192, 384, 96
    hs, ht, ws, wt, cs, ct, cw, trow = _grad2_case(E, K, D, Tr, 9)
    inv_ts, inv_tt = 10.0, 25.0
    d = lambda t: t.to(DEV)
    qt, refs, rb2_t = ops.head_teacher(d(ht), d(wt), inv_tt, d(ct))
    _, lse2_s = ops.head_stats(d(hs), d(ws), inv_ts, d(cs), want_nat=False)
    rb2_e = rb2_t[d(trow)].contiguous()
    l_new = torch.empty(3, device=DEV)
    G, _ = ops.head_grad2(d(hs), d(ws), inv_ts, d(cs), lse2_s, d(cw), rb2_e, d(trow).to(torch.int32), qt, refs, 256, l_new)
    l_old = torch.empty(2, device=DEV)
    ht_e = d(ht)[d(trow)].contiguous()
    Gt, _ = ops.head_grad(d(ws), d(wt), d(hs), ht_e, inv_ts, inv_tt, d(cs), d(ct), None, 256, lse2_s, rb2_e, d(cw), l_old)
    torch.cuda.synchronize()
    assert rel(G, Gt.t()) < 3e-3
    assert abs(l_new[2].item() - l_old.sum().item()) <= 2e-5 * abs(l_old.sum().item())

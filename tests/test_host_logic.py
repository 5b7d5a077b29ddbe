"""Host-side logic that needs no GPU: entry pairing tables of the fused path, workload bookkeeping
(BASELINE.md section 4 numbers), module state-dict contract, loud failure without CUDA."""
import math

import pytest
import torch

from dinox_b200 import losshead, synth
from oracle import losshead_oracle as O


def test_entry_plan_matches_multicrop_definition():
    B, Vg, V, Mm = 3, 2, 5, 7
    plan = losshead._EntryPlan(B, Vg, V, Mm, "cpu")
    es, et, cw = plan.ent_s.tolist(), plan.ent_t.tolist(), plan.cw_base.tolist()
    pairs = {(iq * B + b, v * B + b) for iq in range(Vg) for v in range(V) if v != iq for b in range(B)}
    got = {(t, s) for s, t, w in zip(es[:plan.n_cls], et[:plan.n_cls], cw) if s >= 0}
    assert got == pairs and plan.n_cls == (Vg * V - Vg) * B
    assert plan.e_cls_pad % 128 == 0 and plan.e_pad % 128 == 0
    assert all(abs(w - 1.0 / ((Vg * V - Vg) * B)) < 1e-7 for w in cw[:plan.n_cls])
    assert all(w == 0.0 for w in cw[plan.n_cls:])
    # iBOT entries: 1:1 after the padded CLS block
    for m in range(Mm):
        assert es[plan.e_cls_pad + m] == B * V + m and et[plan.e_cls_pad + m] == B * Vg + m
    # CSR covers every non-padding entry exactly once, grouped by student row
    ptr, ent = plan.csr_ptr.tolist(), plan.csr_ent.tolist()
    assert len(ptr) == B * V + Mm + 1 and sorted(ent) == [e for e, s in enumerate(es) if s >= 0]
    for r in range(B * V + Mm):
        assert all(es[e] == r for e in ent[ptr[r]:ptr[r + 1]])


def test_entry_weights_reproduce_oracle_loss():
    """sum over entries of cw * CE(teacher row, student row) == multicrop_dino_loss (pure torch)."""
    g = torch.Generator().manual_seed(0)
    B, Vg, V, K = 4, 2, 6, 64
    s = torch.randn(B * V, K, generator=g)
    t = torch.randn(B * Vg, K, generator=g)
    c = torch.zeros(1, K)
    ref = O.multicrop_dino_loss(s, t, c, 0.1, 0.04, Vg, V - Vg)
    plan = losshead._EntryPlan(B, Vg, V, 0, "cpu")
    q = torch.softmax(t / 0.04, -1)
    logp = torch.log_softmax(s / 0.1, -1)
    tot = 0.0
    for e in range(plan.n_cls):
        tot += plan.cw_base[e] * -(q[plan.ent_t[e]] * logp[plan.ent_s[e]]).sum()
    assert abs(float(tot) - ref.item()) < 1e-5 * abs(ref.item())


def test_workload_bookkeeping_matches_baseline_md():
    c2 = synth.LossHeadShapes(**synth.CONFIGS["C2"])
    assert (c2.student_rows, c2.teacher_rows, c2.masked_rows, c2.tokens) == (640, 128, 7424, 201)
    assert abs(c2.flops() / 1e9 - 1618.9) < 0.2
    assert abs(c2.hbm_bytes(47_084_800, 4) / 1e6 - 639) < 2
    c1 = synth.LossHeadShapes(**synth.CONFIGS["C1"])
    assert (c1.student_rows, c1.teacher_rows, c1.masked_rows) == (80, 16, 928)
    assert abs(c1.flops() / 1e9 - 202.4) < 0.1
    c4 = synth.LossHeadShapes(**synth.CONFIGS["C4"])
    assert abs(c4.flops() / 1e9 - 2179.3) < 0.5


def test_student_param_shapes_match_survey_appendix_a():
    s = synth.student_param_shapes(384, 12, 65536)
    assert len(s) == 161 and sum(math.prod(x) for x in s) == 47_084_800
    small = sum(1 for x in s if math.prod(x) < 1024)
    assert small == 82 and s[-2] == (65536, 384) and s[1] == (1, 197, 384)
    l = synth.student_param_shapes(1024, 24, 65536)
    assert len(l) == 305 and abs(sum(math.prod(x) for x in l) / 1e6 - 371.8) < 0.1


def test_module_contract_without_gpu():
    dl = losshead.DINOLoss(4096, 0.9)
    assert list(dl.state_dict()) == ["center"] and dl.center.shape == (1, 4096) and dl.center.dtype == torch.float32
    dl.load_state_dict({"center": torch.ones(1, 4096)})
    head = losshead.ProjectionHead(32, 96)
    assert list(head.state_dict()) == ["0.weight", "0.bias", "2.weight", "2.bias"]
    assert head[0].weight.shape == (32, 32) and head[2].weight.shape == (96, 32)

    class BB(torch.nn.Module):
        dim = 32
    m = losshead.DinoStudentTeacher(BB(), out_dim=96)
    assert [k for k in m.state_dict()] == ["head.0.weight", "head.0.bias", "head.2.weight", "head.2.bias"]


def test_cpu_tensors_fail_loudly():
    with pytest.raises(Exception, match="CUDA|fallback"):
        losshead.DINOLoss(16)(torch.zeros(4, 16), torch.zeros(4, 16), 0.1, 0.04)
    with pytest.raises(Exception, match="CUDA|fallback"):
        losshead.compute_gram_anchoring_loss(torch.zeros(1, 5, 8), torch.zeros(1, 5, 8))
    with pytest.raises(Exception, match="CUDA|fallback"):
        losshead.ProjectionHead(8, 16)(torch.zeros(2, 8))
    with pytest.raises(Exception):
        losshead.ema_update([torch.zeros(4)], [torch.zeros(4)], 0.9)


def test_synthetic_ct_crops_are_seeded_and_in_range():
    g1, g2 = synth.seeded_generator(1, 0), synth.seeded_generator(1, 0)
    v1, s1 = synth.multicrop_batch(2, g1, 2, 2, 32, 16)
    v2, s2 = synth.multicrop_batch(2, g2, 2, 2, 32, 16)
    assert all(torch.equal(a, b) for a, b in zip(v1, v2)) and torch.equal(s1, s2)
    assert v1[0].shape == (2, 3, 32, 32) and v1[2].shape == (2, 3, 16, 16) and s1.shape == (2, 3)
    for v in v1:
        assert v.min() >= -2.2 and v.max() <= 2.7
    assert (s1[:, 0] >= 0.46).all() and (s1[:, 2] <= 5.0).all()
    g3 = synth.seeded_generator(1, 1)
    assert not torch.equal(synth.multicrop_batch(2, g3, 2, 2, 32, 16)[0][0], v1[0])


def test_checkpoint_wire_format_matches_reference_modules():
    """the payload of save_checkpoint (scripts/phase5_big_run.py:1104-1125) - student /
    teacher state_dict, dino_loss state_dict, optimizer state_dict - has the same keys, shapes and dtypes
    with the drop-in modules as with plain PyTorch modules of the reference's structure, and round-trips
    through torch.save / load_state_dict(strict=True) in both directions.  (CPU only: no kernels run.)"""
    import io
    import torch
    import torch.nn as nn
    from dinox_b200 import losshead
    from dinox_b200.optim import FusedAdamW

    class Backbone(nn.Module):           # stand-in with the attribute the wrapper needs
        def __init__(self, dim):
            super().__init__()
            self.dim = dim
            self.proj = nn.Linear(dim, dim)

    class RefStudentTeacher(nn.Module):  # structure of zoo/arch.py:246-261
        def __init__(self, backbone, out_dim):
            super().__init__()
            self.backbone = backbone
            self.head = nn.Sequential(nn.Linear(backbone.dim, backbone.dim), nn.GELU(), nn.Linear(backbone.dim, out_dim))

    class RefDINOLoss(nn.Module):        # buffers of scripts/phase5_big_run.py:679-684
        def __init__(self, out_dim):
            super().__init__()
            self.register_buffer("center", torch.zeros(1, out_dim))

    D, K = 32, 96
    ours, ref = losshead.DinoStudentTeacher(Backbone(D), K), RefStudentTeacher(Backbone(D), K)
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys())
    assert all(sd_o[k].shape == sd_r[k].shape and sd_o[k].dtype == sd_r[k].dtype for k in sd_o)
    dl_o, dl_r = losshead.DINOLoss(K, 0.9), RefDINOLoss(K)
    assert {k: (v.shape, v.dtype) for k, v in dl_o.state_dict().items()} == {k: (v.shape, v.dtype) for k, v in dl_r.state_dict().items()}
    opt_o, opt_r = FusedAdamW(ours.parameters(), lr=1e-3, weight_decay=0.04), torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=0.04)
    assert opt_o.state_dict()["param_groups"][0].keys() == opt_r.state_dict()["param_groups"][0].keys()
    # round trip through the on-disk payload in both directions
    buf = io.BytesIO()
    torch.save({"step": 7, "student": sd_r, "teacher": sd_r, "opt": opt_r.state_dict(), "scaler": None,
                "dino_loss": dl_r.state_dict()}, buf)
    buf.seek(0)
    payload = torch.load(buf, weights_only=False)
    ours.load_state_dict(payload["student"], strict=True)
    dl_o.load_state_dict(payload["dino_loss"], strict=True)
    opt_o.load_state_dict(payload["opt"])
    ref.load_state_dict(ours.state_dict(), strict=True)
    dl_r.load_state_dict(dl_o.state_dict(), strict=True)
    assert all(torch.equal(a, b) for a, b in zip(ours.state_dict().values(), ref.state_dict().values()))


def test_patch_index_names_masked_patch_rows_of_the_token_tensor():
    """synth.feature_batch(patches_from_tokens=True): flat row = crop * T + 1 + position, CLS (token 0) and the
    register tokens (last R) are never masked, indices are unique, masks_weight sums to 1 per crop."""
    from dinox_b200 import synth
    sh = synth.LossHeadShapes(batch=3, dim=64, out_dim=256, n_patches=49)
    f = synth.feature_batch(sh, synth.seeded_generator(5, 0), patches_from_tokens=True)
    idx = f["patch_index"]
    assert "student_patch" not in f and idx.dtype == torch.int64 and idx.numel() == sh.masked_rows
    assert idx.unique().numel() == idx.numel()
    tok = idx % sh.tokens
    assert int(tok.min()) >= 1 and int(tok.max()) <= sh.n_patches
    crops = idx // sh.tokens
    assert torch.equal(crops.bincount(minlength=sh.teacher_rows), torch.full((sh.teacher_rows,), sh.masked_per_crop))
    assert abs(float(f["masks_weight"].sum()) - sh.teacher_rows) < 1e-5


def test_shard_layout_of_the_sharded_optimizer():
    """SURVEY 8f #2 (second half): only tensors >= 1 Mi elements whose rows divide by the world size get a
    sharded AdamW state (the head's W2); everything else stays replicated; world 1 shards nothing."""
    from dinox_b200.optim import shard_layout
    head = [(384, 384), (384,), (65536, 384), (65536,)]
    assert shard_layout(head, 8) == [(False, None), (False, None), (True, 8192), (False, None)]
    assert shard_layout(head, 1) == [(False, None)] * 4
    assert shard_layout([(65537, 384)], 8) == [(False, None)]
    assert shard_layout([(1024, 1024)], 4, min_numel=1 << 20) == [(True, 256)]


def test_entry_plan_padded_teacher_layout_of_the_readback_path():
    """Teacher rows of the read-back path: [CLS rows | zero rows to a multiple of 128 | masked patch rows]; `trow`
    names each entry's teacher row in that layout, padding entries point at row 0 with weight 0."""
    B, Vg, V, Mm = 5, 2, 4, 9
    plan = losshead._EntryPlan(B, Vg, V, Mm, "cpu")
    Mt = B * Vg
    assert plan.Mt_pad == 128 and plan.cls_rows_pad.tolist() == list(range(Mt)) + [-1] * (128 - Mt)
    et, etp, trow = plan.ent_t.tolist(), plan.ent_t_pad.tolist(), plan.trow.tolist()
    for e, (a, b, c) in enumerate(zip(et, etp, trow)):
        if a < 0:
            assert b == -1 and c == 0 and plan.cw_base[e] == 0.0
        elif a < Mt:
            assert b == a == c
        else:
            assert b == a - Mt + plan.Mt_pad == c
    assert plan.row_counts.tolist()[:2] == [float(Mt), float(Mm)]
    # without iBOT rows there is nothing to pad
    assert losshead._EntryPlan(B, Vg, V, 0, "cpu").Mt_pad == Mt


def test_precision_and_cache_switches_validate_their_arguments():
    prev = losshead.set_contraction_precision("fp32")
    assert losshead.set_contraction_precision(prev) == "fp32"
    with pytest.raises(ValueError):
        losshead.set_contraction_precision("tf32")
    with losshead.contraction_precision("bf16"):
        assert not losshead._fp32_mode()
    prev = losshead.set_weight_cache("tracked")
    assert losshead.set_weight_cache(prev) == "tracked"
    with pytest.raises(ValueError):
        losshead.set_weight_cache("sometimes")
    with pytest.raises(ValueError):
        losshead.DINOLoss(64, patch_teacher_mode="mean")


def test_balanced_schedule_plan_and_item_walk():
    """Host side of dinox_gemm_bf16_balanced (no GPU): the plan for the step's two big backward GEMMs, and for a range
    of tile / cluster counts the item list the clusters walk - every whole tile once, every part of every cut tile
    once, and the predecessor of a part (same tile, part - 1) always has a LOWER item number (what makes the ordered
    accumulation deadlock free on a resident grid) and lives in the same or an earlier round."""
    import ctypes
    import numpy as np
    from dinox_b200 import _ext
    lib = _ext.lib()

    def plan(tiles, clusters, kblocks):
        first, parts = ctypes.c_int(-1), ctypes.c_int(-1)
        assert lib.dinox_plan_ordered_split(tiles, clusters, kblocks, ctypes.byref(first), ctypes.byref(parts)) == 0
        return first.value, parts.value

    # dW2 at C2: 256 tile pairs, 74 CTA pairs, E = 8192 entries = 128 k-blocks -> the 34 tiles of the last wave in two
    assert plan(256, 74, 128) == (222, 2)
    # dH at C2: 32 tile pairs, K = 65536 prototypes = 1024 k-blocks -> every tile in 9 parts (288 items = 3.9 rounds)
    assert plan(32, 74, 1024) == (0, 9)
    assert plan(74, 74, 64)[1] == 0 and plan(148, 74, 64)[1] == 0          # whole waves: nothing to cut
    assert plan(75, 74, 16)[1] == 0                                          # parts would be shorter than 32 k-blocks

    for num_m, num_n, m_fastest, clusters, kblocks in [(256, 1, 0, 74, 128), (32, 1, 1, 74, 1024), (12, 3, 1, 74, 256),
                                                       (100, 2, 0, 148, 96), (5, 1, 1, 7, 640), (1, 1, 1, 148, 128)]:
        tiles = num_m * num_n
        first, parts = plan(tiles, clusters, kblocks)
        if parts < 2:
            continue
        total = first + (tiles - first) * parts
        out = np.full((total + 8, 6), -7, dtype=np.int32)
        n = lib.dinox_debug_walk_ordered(num_m, num_n, m_fastest, first, parts, clusters, 2,
                                         out.ctypes.data_as(ctypes.c_void_p), out.shape[0])
        assert n == total
        items = out[:n]
        assert sorted(items[:, 0].tolist()) == list(range(total))            # every item number exactly once
        assert (items[:, 1] == items[:, 0] % clusters).all()                 # round-robin over the clusters
        seen = {}
        for item, cid, m_tile, n_tile, kpart, kparts in items.tolist():
            assert m_tile % 2 == 0 and 0 <= m_tile // 2 < num_m and 0 <= n_tile < num_n
            key = (m_tile, n_tile)
            seen.setdefault(key, {})[kpart] = (item, kparts)
        assert len(seen) == tiles
        n_whole = 0
        for key, ps in seen.items():
            kparts = next(iter(ps.values()))[1]
            assert sorted(ps) == list(range(kparts)) and all(v[1] == kparts for v in ps.values())
            n_whole += kparts == 1
            for j in range(1, kparts):
                assert ps[j - 1][0] < ps[j][0]                               # predecessor has a lower item number
                assert ps[j - 1][0] // clusters <= ps[j][0] // clusters      # ... same or earlier round
        assert n_whole == first


def test_loss_term_seeds_round_like_the_fanout_kernel():
    """LossHeadStep seeds term i's backward with fp32((1 * 1/accum) * w_i) - the arithmetic of the scalar_fanout kernel of
    combine_losses - so the decoupled and the summed backward agree bit for bit."""
    import numpy as np
    from dinox_b200.step import LossHeadStep
    st = LossHeadStep.__new__(LossHeadStep)
    st._seeds, st.device = {}, "cpu"
    for accum in (1, 2, 3, 4, 7):
        st.accum = accum
        st._seeds.clear()
        for w in (1.0, 0.5, 0.1, 3.0):
            v = np.float32(np.float32(1.0) * np.float32(1.0 / accum)) * np.float32(w)
            assert st._seed(w).item() == float(v)


class _CeStandIns:
    """torch stand-ins for the row kernels `DINOLoss.forward` calls, recording which ones ran."""

    def __init__(self, max_views=12):
        self.calls, self.max_views = [], max_views

    def ce_onepass_max_views(self):
        return self.max_views

    def axpb(self, a, alpha, beta=0.0, out=None):
        return a * alpha + beta

    def rows_lse(self, x, inv_tau, colbias=None, want_entropy=False):
        self.calls.append("rows_lse")
        u = x.float() * inv_tau - (colbias[None, :] if colbias is not None else 0.0)
        return torch.logsumexp(u, dim=1)

    def cols_lse(self, x, inv_tau, rowbias=None):
        self.calls.append("cols_lse")
        u = x.float() * inv_tau - (rowbias[:, None] if rowbias is not None else 0.0)
        return torch.logsumexp(u, dim=0)

    def lse_combine(self, gathered, add=0.0):
        return torch.logsumexp(gathered, dim=0) + add

    def _loss(self, s, t, B, V, Vg, inv_ts, inv_tt, colb, rowb, lse_s, norm):
        q = torch.exp(t.float() * inv_tt - colb[None, :] - rowb[:, None])
        logp = s.float() * inv_ts - lse_s[:, None]
        tot = 0.0
        for iq in range(Vg):
            for v in range(V):
                if v != iq:
                    tot = tot - (q[iq * B:(iq + 1) * B] * logp[v * B:(v + 1) * B]).sum()
        return tot * norm

    def ce_fwd(self, s, t, B, V, Vg, inv_ts, inv_tt, colb, rowb, lse_s, group_w, norm, exclude_same):
        self.calls.append("ce_fwd")
        return self._loss(s, t, B, V, Vg, inv_ts, inv_tt, colb, rowb, lse_s, norm)

    def ce_fwd_onepass(self, s, t, B, V, Vg, inv_ts, inv_tt, colb, group_w, norm, exclude_same):
        self.calls.append("ce_fwd_onepass")
        lse_s = torch.logsumexp(s.float() * inv_ts, dim=1)
        rowb = torch.logsumexp(t.float() * inv_tt - colb[None, :], dim=1)
        return self._loss(s, t, B, V, Vg, inv_ts, inv_tt, colb, rowb, lse_s, norm), lse_s, rowb

    def cols_sum(self, x, out=None):
        return x.float().sum(0)

    def center_ema_(self, center, colsum, global_rows, momentum):
        center.copy_(center * momentum + (colsum / global_rows).reshape(1, -1) * (1 - momentum))


@pytest.mark.parametrize("mode,teacher,V,expect", [
    ("onepass", "center", 2, ["ce_fwd_onepass"]),                               # the reference's call: one pass
    ("onepass", "center", 10, ["ce_fwd_onepass"]),
    ("onepass", "center", 14, ["rows_lse", "rows_lse", "ce_fwd"]),              # more views than the kernel holds
    ("passes", "center", 2, ["rows_lse", "rows_lse", "ce_fwd"]),
    ("onepass", "sinkhorn", 2, ["cols_lse", "rows_lse"] * 3 + ["rows_lse", "ce_fwd"]),   # SK row offsets are inputs
])
def test_dinoloss_forward_dispatch_and_value(monkeypatch, mode, teacher, V, expect):
    """Which kernels the drop-in forward takes (one pass vs teacher LSE + student LSE + CE), on CPU with torch
    stand-ins for the kernels; every route returns the oracle's loss and (centre mode) moves the centre alike."""
    k = _CeStandIns()
    monkeypatch.setattr(losshead, "ops", k)
    monkeypatch.setattr(losshead, "_as_rows", lambda t: t)
    monkeypatch.setattr(losshead._RowCE, "apply", staticmethod(lambda s, t, cb, rb, lse, gw, B, V_, Vg, its, itt, n, ex:
                                                               k.ce_fwd(s, t, B, V_, Vg, its, itt, cb, rb, lse, gw, n, ex)))
    monkeypatch.setattr(losshead._RowCEOnePass, "apply", staticmethod(lambda s, t, cb, gw, B, V_, Vg, its, itt, n, ex:
                                                                      k.ce_fwd_onepass(s, t, B, V_, Vg, its, itt, cb, gw, n, ex)[0]))
    monkeypatch.setattr(losshead, "sinkhorn_knopp_biases",
                        (lambda f: (lambda t, tau, n=3, pg=None, _k=None: f(t, tau, n, pg, _k=k)))(losshead.sinkhorn_knopp_biases))
    orig_update = losshead.DINOLoss.update_center      # its kernel namespace is a default argument
    monkeypatch.setattr(losshead.DINOLoss, "update_center", lambda self, t: orig_update(self, t, _k=k))
    prev = losshead.set_ce_forward(mode)
    try:
        g = torch.Generator().manual_seed(5)
        B, Vg, K = 3, 2, 96
        s = torch.randn(V * B, K, generator=g)
        t = torch.randn(Vg * B, K, generator=g)
        crit = losshead.DINOLoss(K, 0.9, n_global=Vg, n_local=V - Vg, teacher_mode=teacher)
        c0 = torch.randn(1, K, generator=g) * 0.1
        crit.center.copy_(c0)
        loss = crit(s, t, 0.1, 0.04)
    finally:
        losshead.set_ce_forward(prev)
    assert k.calls == expect, k.calls
    ref = O.multicrop_dino_loss(s, t, c0, 0.1, 0.04, Vg, V - Vg, teacher_mode=teacher)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    if teacher == "center":
        assert torch.allclose(crit.center, c0 * 0.9 + t.mean(0, keepdim=True) * 0.1, atol=1e-6)
    else:
        assert torch.equal(crit.center, c0)


def test_ce_forward_switch_validates():
    with pytest.raises(ValueError):
        losshead.set_ce_forward("twice")
    assert losshead.set_ce_forward("onepass") in ("onepass", "passes")

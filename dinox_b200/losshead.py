"""Host-side mirror of the reference's loss-head API, backed by the CUDA kernels.

Drop-in names and signatures (reference: timlawrenz/DINO-X):
  DINOLoss(out_dim, center_momentum)(student_out, teacher_out, student_temp, teacher_temp)
                                              scripts/phase5_big_run.py:679-720
  compute_gram_matrix(feats)                  scripts/phase5_big_run.py:723-728
  compute_gram_anchoring_loss(s_feats, t_feats)  scripts/phase5_big_run.py:731-739
  DinoStudentTeacher(backbone, out_dim).head  zoo/arch.py:246-261 (state-dict keys head.{0,2}.*)
  _ema_update(teacher, student, m)            scripts/phase3_micro_run.py:152-155; inline loop at
                                              scripts/phase5_big_run.py:1798-1802
  KoLeoLoss()(student_out)                   scripts/phase5_big_run.py:742-773
Extensions enter only through keyword arguments whose defaults reproduce the reference
(n_global=2, n_local=0, teacher_mode="center", process_group=None) and through the fused entry
point ``fused_head_dino_loss`` (projection head + multi-crop CE + iBOT in tcgen05 GEMM epilogues,
logits never written to HBM).

Everything computes through libdinox_b200.so; CPU tensors raise (no fallback).
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _ext, ops

LOG2E = ops.LOG2E


# =================================================================================================
# data-parallel statistics: only tiny all-reduces of global statistics
# =================================================================================================
def _world(pg) -> int:
    if pg is None:
        return 1
    import torch.distributed as dist
    return dist.get_world_size(pg) if pg is not True else dist.get_world_size()


def _group(pg):
    return None if pg is True else pg


def allreduce_sum_(t: torch.Tensor, pg) -> torch.Tensor:
    """In-place SUM all-reduce over the data-parallel group (NCCL over NVLink on B200)."""
    if _world(pg) > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group(pg))
    return t


def allreduce_sum_async(t: torch.Tensor, pg):
    """SUM all-reduce that returns a work handle (None when single-process); the caller waits on it
    right before the result is consumed, so the collective overlaps the kernels queued in between."""
    if _world(pg) > 1:
        import torch.distributed as dist
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group(pg), async_op=True)
    return None


def allreduce_lse(local_lse: torch.Tensor, pg, add: float = 0.0, _k=ops) -> torch.Tensor:
    """log-sum-exp all-reduce of per-rank LSE vectors: all_gather + one combine kernel.  Exact in
    the log domain (no common shift needed), one collective per Sinkhorn half-iteration."""
    w = _world(pg)
    if w == 1:
        return _k.lse_combine(local_lse.reshape(1, -1), add) if add != 0.0 else local_lse
    import torch.distributed as dist
    gathered = [torch.empty_like(local_lse) for _ in range(w)]
    dist.all_gather(gathered, local_lse.contiguous(), group=_group(pg))
    return _k.lse_combine(torch.stack(gathered, 0), add)


# =================================================================================================
# E2 Sinkhorn-Knopp on materialised teacher logits (small: CLS rows only)
# =================================================================================================
@torch.no_grad()
def sinkhorn_knopp_biases(teacher_out: torch.Tensor, teacher_temp: float, n_iterations: int = 3,
                          process_group=None, _k=ops) -> Tuple[torch.Tensor, torch.Tensor]:
    """Log-domain Sinkhorn-Knopp (DINOv2 formulation, see oracle.sinkhorn_knopp).  Returns
    (colbias a[K], rowbias b[rows]) with q[i,k] = exp(t[i,k]/tau - a[k] - b[i]); rows of q sum to 1.
    Per-prototype sums are all-reduced over `process_group`; per-sample sums are local."""
    rows, K = teacher_out.shape
    w = _world(process_group)
    inv_tau = 1.0 / teacher_temp
    log_bg, log_k = math.log(rows * w), math.log(K)
    if n_iterations < 1:
        raise ValueError("Sinkhorn-Knopp needs at least one iteration")
    b = None
    a = None
    # The scalar normalisations (Q /= K after the prototype step, Q /= B after the sample step, the final Q *= B) are
    # shifts of a / b in the log domain.  In a single process they ride in the one axpb per iteration instead of a
    # combine launch of their own (a_k -> a_k + c shifts b_i by -c), and the last iteration's "/= B" cancels against
    # the final "*= B": 10 launches for three iterations instead of 13.  Across ranks the combine kernel behind the
    # all-gather adds log K for free.
    for it in range(n_iterations):
        a_local = _k.cols_lse(teacher_out, inv_tau, b)             # LSE_i(x - b_i) over local samples
        if w == 1:
            a, pending = a_local, log_k                             # true a = a_local + log K, applied below
        else:
            a, pending = allreduce_lse(a_local, process_group, add=log_k, _k=_k), 0.0
        b = _k.rows_lse(teacher_out, inv_tau, a)                    # LSE_k(x - a_k) (+ pending)
        last = it == n_iterations - 1
        shift = (0.0 if last else log_bg) - pending                 # Q /= B, undone again after the last iteration
        if shift != 0.0:
            b = _k.axpb(b, 1.0, shift)
    if w == 1:
        a = _k.axpb(a, 1.0, log_k)
    return a, b


@torch.no_grad()
def sinkhorn_knopp_teacher(teacher_out: torch.Tensor, teacher_temp: float, n_iterations: int = 3,
                           process_group=None) -> torch.Tensor:
    """Materialised SK assignment (rows, K) fp32 - convenience for tests / diagnostics."""
    a, b = sinkhorn_knopp_biases(teacher_out, teacher_temp, n_iterations, process_group)
    # q = exp(t/tau - a - b): reuse the CE backward kernel algebra is overkill; this is a debug helper
    return torch.exp(teacher_out.float() / teacher_temp - a[None, :] - b[:, None])


# =================================================================================================
# a2-a5, E1: DINOLoss on materialised logits
# =================================================================================================
class _RowCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, colbias_t, rowbias_t, lse_s, group_w, groups, V, Vg, inv_ts, inv_tt,
                norm, exclude_same):
        loss = ops.ce_fwd(student, teacher, groups, V, Vg, inv_ts, inv_tt, colbias_t, rowbias_t, lse_s, group_w,
                          norm, exclude_same)
        ctx.save_for_backward(student, teacher, colbias_t, rowbias_t, lse_s, group_w)
        ctx.cfg = (groups, V, Vg, inv_ts, inv_tt, norm, exclude_same)
        return loss

    @staticmethod
    def backward(ctx, g):
        student, teacher, colbias_t, rowbias_t, lse_s, group_w = ctx.saved_tensors
        groups, V, Vg, inv_ts, inv_tt, norm, exclude_same = ctx.cfg
        grad = ops.ce_bwd(student, teacher, groups, V, Vg, inv_ts, inv_tt, colbias_t, rowbias_t, lse_s, group_w,
                          norm, exclude_same, g)
        return (grad,) + (None,) * 12


class _RowCEOnePass(torch.autograd.Function):
    """Softmax-centred teacher: loss, student LSEs and teacher LSEs from ONE pass over the logits
    (`dinox_ce_fwd_onepass`); the backward is the same kernel as `_RowCE`."""

    @staticmethod
    def forward(ctx, student, teacher, colbias_t, group_w, groups, V, Vg, inv_ts, inv_tt, norm, exclude_same):
        loss, lse_s, rowbias_t = ops.ce_fwd_onepass(student, teacher, groups, V, Vg, inv_ts, inv_tt, colbias_t,
                                                    group_w, norm, exclude_same)
        ctx.save_for_backward(student, teacher, colbias_t, rowbias_t, lse_s, group_w)
        ctx.cfg = (groups, V, Vg, inv_ts, inv_tt, norm, exclude_same)
        return loss

    @staticmethod
    def backward(ctx, g):
        student, teacher, colbias_t, rowbias_t, lse_s, group_w = ctx.saved_tensors
        groups, V, Vg, inv_ts, inv_tt, norm, exclude_same = ctx.cfg
        grad = ops.ce_bwd(student, teacher, groups, V, Vg, inv_ts, inv_tt, colbias_t, rowbias_t, lse_s, group_w,
                          norm, exclude_same, g)
        return (grad,) + (None,) * 10


# DINOLoss.forward on materialised logits with the softmax-centred teacher: "onepass" (default) reads every logit once
# in the forward (dinox_ce_fwd_onepass); "passes" is the three-pass form (teacher LSE, student LSE, cross-entropy)
# that the Sinkhorn-Knopp teacher and V > 12 views always take.  DINOX_CE_FORWARD overrides at import.
_CE_FORWARD = [os.environ.get("DINOX_CE_FORWARD", "onepass")]
if _CE_FORWARD[0] not in ("onepass", "passes"):
    raise ValueError(f"DINOX_CE_FORWARD={_CE_FORWARD[0]!r}: expected onepass or passes")


def set_ce_forward(mode: str) -> str:
    """Selects the forward of `DINOLoss` on materialised logits ("onepass" or "passes"); returns the previous mode."""
    if mode not in ("onepass", "passes"):
        raise ValueError(f"ce forward mode {mode!r}: expected 'onepass' or 'passes'")
    prev = _CE_FORWARD[0]
    _CE_FORWARD[0] = mode
    return prev


def _as_rows(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise _ext.DinoxError("dinox_b200: CUDA tensors required (no CPU fallback)")
    if t.dim() != 2:
        raise ValueError(f"expected (rows, K) logits, got {tuple(t.shape)}")
    return t if t.stride(1) == 1 else t.contiguous()


class DINOLoss(nn.Module):
    """DINO loss with centering and sharpening - drop-in for scripts/phase5_big_run.py:679-720.

    Extra keyword arguments (all defaulting to the reference behaviour):
      n_global / n_local : multi-crop layout of `student_out` (view-major rows), E1
      teacher_mode       : "center" (reference) or "sinkhorn" (E2; center buffer untouched)
      process_group      : data-parallel group for the centre / Sinkhorn statistics (True = WORLD)
    """

    def __init__(self, out_dim: int, center_momentum: float = 0.999, *, n_global: int = 2, n_local: int = 0,
                 teacher_mode: str = "center", sk_iterations: int = 3, process_group=None,
                 patch_teacher_mode: Optional[str] = None) -> None:
        super().__init__()
        self.center_momentum = center_momentum
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.n_global, self.n_local = n_global, n_local
        self.teacher_mode, self.sk_iterations = teacher_mode, sk_iterations
        # iBOT (masked-patch) teacher rows of the fused path: "center" (default: softmax-centred with the patch
        # centre, which is kept updated in both CLS modes) or "sinkhorn" (DINOv2: Sinkhorn-Knopp over the masked
        # patches of the global batch as well - their teacher logits are materialised for the three iterations)
        self.patch_teacher_mode = patch_teacher_mode or "center"
        if self.patch_teacher_mode not in ("center", "sinkhorn"):
            raise ValueError(f"unknown patch_teacher_mode {patch_teacher_mode!r}")
        self.process_group = process_group

    @torch.no_grad()
    def update_center(self, teacher_output: torch.Tensor, _k=ops) -> None:
        """center <- center*m + mean_rows(teacher_output)*(1-m)  (:686-690); in DP the column sums are
        all-reduced and divided by the global row count."""
        t = _as_rows(teacher_output) if _k is ops else teacher_output
        colsum = _k.cols_sum(t)
        allreduce_sum_(colsum, self.process_group)
        _k.center_ema_(self.center, colsum, t.shape[0] * _world(self.process_group), self.center_momentum)

    def teacher_biases(self, teacher_out: torch.Tensor, teacher_temp: float):
        if self.teacher_mode == "center":
            colbias = ops.axpb(self.center.reshape(-1), 1.0 / teacher_temp)
            rowbias = ops.rows_lse(teacher_out, 1.0 / teacher_temp, colbias)
            return colbias, rowbias
        if self.teacher_mode == "sinkhorn":
            return sinkhorn_knopp_biases(teacher_out, teacher_temp, self.sk_iterations, self.process_group)
        raise ValueError(f"unknown teacher_mode {self.teacher_mode!r}")

    def forward(self, student_out: torch.Tensor, teacher_out: torch.Tensor, student_temp: float,
                teacher_temp: float) -> torch.Tensor:
        s, t = _as_rows(student_out), _as_rows(teacher_out.detach())
        Vg = self.n_global
        if t.shape[0] % Vg:
            raise ValueError("teacher rows must be a multiple of n_global")
        B = t.shape[0] // Vg
        if s.shape[0] % B:
            raise ValueError("student rows must be a multiple of the per-view batch")
        V = s.shape[0] // B
        if self.n_local and V != Vg + self.n_local:
            raise ValueError(f"student rows imply {V} views, expected {Vg + self.n_local}")
        n_terms = Vg * V - Vg
        onepass = (self.teacher_mode == "center" and _CE_FORWARD[0] == "onepass" and V <= ops.ce_onepass_max_views()
                   and s.shape[0] * s.stride(0) < 2 ** 32 and t.shape[0] * t.stride(0) < 2 ** 32)   # 32-bit offsets
        if onepass:
            with torch.no_grad():
                colbias = ops.axpb(self.center.reshape(-1), 1.0 / teacher_temp)
            loss = _RowCEOnePass.apply(s, t, colbias, None, B, V, Vg, 1.0 / student_temp, 1.0 / teacher_temp,
                                       1.0 / (n_terms * B), True)
        else:
            with torch.no_grad():
                colbias, rowbias = self.teacher_biases(t, teacher_temp)
                lse_s = ops.rows_lse(s.detach(), 1.0 / student_temp)
            loss = _RowCE.apply(s, t, colbias, rowbias, lse_s, None, B, V, Vg, 1.0 / student_temp, 1.0 / teacher_temp,
                                1.0 / (n_terms * B), True)
        if self.teacher_mode == "center":
            self.update_center(t)
        return loss


@torch.no_grad()
def entropy_diagnostics(student_out, teacher_out, center, student_temp, teacher_temp):
    """Teacher / student softmax entropies (scripts/phase5_big_run.py:1843-1853) from one LSE pass each."""
    t, s = _as_rows(teacher_out), _as_rows(student_out.detach())
    colbias = ops.axpb(center.reshape(-1).float(), 1.0 / teacher_temp)
    _, t_ent = ops.rows_lse(t, 1.0 / teacher_temp, colbias, want_entropy=True)
    _, s_ent = ops.rows_lse(s, 1.0 / student_temp, None, want_entropy=True)
    return t_ent.mean(), s_ent.mean()


# =================================================================================================
# a6-a7 Gram anchoring
# =================================================================================================
class _GramMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats):
        xn, inv = ops.normalize_tokens(feats, skip=0)
        gram = ops.gemm_bf16_batched(xn, xn)
        ctx.save_for_backward(feats, xn, inv)
        return gram

    @staticmethod
    def backward(ctx, dg):
        feats, xn, inv = ctx.saved_tensors
        Bt, T, D = feats.shape
        ld = (T + 7) // 8 * 8
        dgb = torch.zeros(Bt, T, ld, dtype=torch.bfloat16, device=dg.device)
        dgb[:, :, :T] = dg  # cast; rare path (the training loop uses compute_gram_anchoring_loss)
        dgv = dgb[:, :, :T]
        dxn = ops.gemm_bf16_batched(dgv, xn, b_mn_major=True)                       # dG   @ Xn
        ops.gemm_bf16_batched(dgv, xn, a_mn_major=True, b_mn_major=True, out=dxn, accumulate=True)  # dG^T @ Xn
        grad = torch.empty(Bt, T, D, dtype=torch.float32, device=dg.device)
        ops.normalize_tokens_bwd(feats, dxn, inv, grad, skip=0)
        return grad.to(feats.dtype)


def compute_gram_matrix(feats: torch.Tensor) -> torch.Tensor:
    """Gram matrix of L2-normalised tokens, (B, N, D) -> (B, N, N)  (scripts/phase5_big_run.py:723-728)."""
    if not feats.is_cuda:
        raise _ext.DinoxError("dinox_b200: CUDA tensors required (no CPU fallback)")
    return _GramMatrixFp32.apply(feats) if _fp32_mode() else _GramMatrix.apply(feats)


class _GramAnchor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student_feats, teacher_feats):
        Bt, T, D = student_feats.shape
        xs, inv_s = ops.normalize_tokens(student_feats, skip=1)
        xt, _ = ops.normalize_tokens(teacher_feats, skip=1)
        n = Bt * (T - 1) * (T - 1)
        need_grad = student_feats.requires_grad
        with ops.TIMER.region("gram_diff"):
            loss, delta = ops.gram_diff(xs, xt, 1.0 / n, want_delta=need_grad)
        if need_grad:
            ctx.save_for_backward(student_feats, xs, inv_s, delta)
            ctx.n = n
        return loss

    @staticmethod
    def backward(ctx, g):
        student_feats, xs, inv_s, delta = ctx.saved_tensors
        Bt, T, D = student_feats.shape
        dv = delta[:, :, :T - 1]
        # dL/dXn = (dG + dG^T) Xn = (4/n) * Delta @ Xn   (Delta symmetric)
        dxn = ops.gemm_bf16_batched(dv, xs, b_mn_major=True, alpha=4.0 / ctx.n)
        # the CLS row receives no gradient (the reference slices feats[:, 1:]): the kernel writes it as zeros
        grad = torch.empty(Bt, T, D, dtype=torch.float32, device=g.device)
        up = g.to(torch.float32).reshape(1).contiguous()
        ops.normalize_tokens_bwd(student_feats, dxn, inv_s, grad, skip=1, scale_dev=up)
        return grad.to(student_feats.dtype), None


def compute_gram_anchoring_loss(student_feats: torch.Tensor, teacher_feats: torch.Tensor) -> torch.Tensor:
    """mse_loss(Gram(student[:,1:]), Gram(teacher[:,1:]))  (scripts/phase5_big_run.py:731-739); the two
    Gram matrices live only in TMEM, only their bf16 difference is kept for the backward GEMM."""
    if not (student_feats.is_cuda and teacher_feats.is_cuda):
        raise _ext.DinoxError("dinox_b200: CUDA tensors required (no CPU fallback)")
    if student_feats.shape != teacher_feats.shape or student_feats.dim() != 3:
        raise ValueError("student/teacher feats must both be (B, T, D)")
    fn = _GramAnchorFp32 if _fp32_mode() else _GramAnchor
    return fn.apply(student_feats, teacher_feats.detach())


# =================================================================================================
# a8 EMA
# =================================================================================================
_EMA_PLANS: Dict[int, ops.EmaPlan] = {}


@torch.no_grad()
def ema_update(teacher_params: Sequence[torch.Tensor], student_params: Sequence[torch.Tensor], m: float,
               plan_key: Optional[int] = None) -> None:
    """p_t <- m*p_t + (1-m)*p_s for every pair, ONE kernel launch (scripts/phase5_big_run.py:1798-1802)."""
    tp = [p.data for p in teacher_params]
    sp = [p.data for p in student_params]
    key = plan_key if plan_key is not None else hash(tuple(t.data_ptr() for t in tp))
    plan = _EMA_PLANS.get(key)
    if plan is None or not plan.matches(sp, tp):
        plan = ops.EmaPlan(sp, tp)
        _EMA_PLANS[key] = plan
    with ops.TIMER.region("ema_multi"):
        plan.apply(m)
    _WEIGHT_EPOCH[0] += 1


def _ema_update(teacher: nn.Module, student: nn.Module, m: float) -> None:
    """Same name/arguments as scripts/phase3_micro_run.py:152-155."""
    ema_update(list(teacher.parameters()), list(student.parameters()), m, plan_key=id(teacher))


# =================================================================================================
# a1 projection head as an nn.Sequential with the reference's parameter keys
# =================================================================================================
_BF16_CACHE: Dict[int, Tuple] = {}
_WEIGHT_EPOCH = [0]  # bumped by every dinox kernel that writes parameters through raw pointers
_WEIGHT_CACHE_MODE = ["always"]


def set_weight_cache(mode: str) -> str:
    """How the bf16 operand copies of the head weights are kept:

    "always"  (default) - re-cast from the fp32 master on every forward, exactly like CUDA autocast does.  Safe
              for ANY way of updating parameters, including the reference's own EMA loop
              `p_t.data.mul_(m).add_(p_s.data, alpha=1-m)` (scripts/phase5_big_run.py:1800-1802), which does
              not bump `p._version`.
    "tracked" - reuse the copy until the parameter's version counter / storage changes or a dinox kernel that
              writes parameters (`ema_update`, `FusedAdamW.step`, `ShardedFusedAdamW.step`) bumps the weight
              epoch.  For loops that update parameters ONLY through autograd-visible in-place ops or those
              dinox entry points (`LossHeadStep` does); call `invalidate_weight_cache()` after anything else.
    Returns the previous mode."""
    if mode not in ("always", "tracked"):
        raise ValueError(f"weight cache mode {mode!r}: expected 'always' or 'tracked'")
    prev = _WEIGHT_CACHE_MODE[0]
    _WEIGHT_CACHE_MODE[0] = mode
    return prev


def invalidate_weight_cache() -> None:
    """Forces a re-cast of every cached bf16 weight copy at its next use (copies are rewritten in place, so
    captured CUDA graphs keep valid addresses)."""
    _WEIGHT_EPOCH[0] += 1


def bf16_weight(p: torch.Tensor) -> torch.Tensor:
    """bf16 copy of a weight.  In "tracked" mode an entry is valid only for the very same tensor object (weak
    reference: Python ids and allocator blocks are both recycled), the same version counter, storage pointer
    and weight epoch; in "always" mode it is rewritten on every call.  The copy of a live parameter is always
    refreshed IN PLACE: captured CUDA graphs hold its address."""
    import weakref
    key = id(p)
    ver = p._version
    hit = _BF16_CACHE.get(key)
    same = hit is not None and hit[0]() is p and hit[4].shape == p.shape and hit[4].device == p.device
    if (same and _WEIGHT_CACHE_MODE[0] == "tracked" and hit[1] == ver and hit[2] == p.data_ptr()
            and hit[3] == _WEIGHT_EPOCH[0]):
        return hit[4]
    w = p.detach()
    out = hit[4] if same else torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
    ops.gather_cast_bf16(w.reshape(w.shape[0], -1), None, out.reshape(w.shape[0], -1))
    if len(_BF16_CACHE) > 64:   # drop entries whose parameter is gone
        for k in [k for k, v in _BF16_CACHE.items() if v[0]() is None]:
            del _BF16_CACHE[k]
    _BF16_CACHE[key] = (weakref.ref(p), ver, p.data_ptr(), _WEIGHT_EPOCH[0], out)
    return out


def _to_bf16_rows(x: torch.Tensor) -> torch.Tensor:
    x = x if x.stride(-1) == 1 else x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    ops.gather_cast_bf16(x, None, out)
    return out


# -------------------------------------------------------------------------------------------------
# Contraction precision of the drop-in modules
# -------------------------------------------------------------------------------------------------
_PRECISION = ["auto"]


def set_contraction_precision(mode: str) -> str:
    """Precision of the dense contractions behind `ProjectionHead`, `compute_gram_matrix` and
    `compute_gram_anchoring_loss`:

    "auto" (default) - what the reference does on the same call: inside `torch.autocast` (its `--amp` flag) bf16
             tensor-core operands with fp32 accumulation; OUTSIDE autocast (the reference's default,
             scripts/phase5_big_run.py:1322) fp32-faithful products - three bf16 GEMMs on hi/lo operand splits, ~16
             mantissa bits per product (see csrc/precise.cu).
    "bf16" - always bf16 operands (fastest; what `fused_head_dino_loss` always uses).
    "fp32" - always the fp32-faithful products.
    Returns the previous mode."""
    if mode not in ("auto", "bf16", "fp32"):
        raise ValueError(f"contraction precision {mode!r}: expected 'auto', 'bf16' or 'fp32'")
    prev = _PRECISION[0]
    _PRECISION[0] = mode
    return prev


@contextlib.contextmanager
def contraction_precision(mode: str):
    """`with contraction_precision("bf16"): ...` - scoped form of set_contraction_precision."""
    prev = set_contraction_precision(mode)
    try:
        yield
    finally:
        _PRECISION[0] = prev


def _fp32_mode() -> bool:
    m = _PRECISION[0]
    return m == "fp32" or (m == "auto" and not torch.is_autocast_enabled())


_SPLIT_CACHE: Dict[int, Tuple] = {}


def split_weight(p: torch.Tensor):
    """(hi, lo) bf16 split of a weight, cached like `bf16_weight` (same invalidation rules, "tracked" mode only)."""
    import weakref
    key = id(p)
    hit = _SPLIT_CACHE.get(key)
    if (hit is not None and hit[0]() is p and _WEIGHT_CACHE_MODE[0] == "tracked" and hit[1] == p._version
            and hit[2] == p.data_ptr() and hit[3] == _WEIGHT_EPOCH[0]):
        return hit[4]
    pair = ops.split_bf16(p.detach())
    if len(_SPLIT_CACHE) > 64:
        for k in [k for k, v in _SPLIT_CACHE.items() if v[0]() is None]:
            del _SPLIT_CACHE[k]
    _SPLIT_CACHE[key] = (weakref.ref(p), p._version, p.data_ptr(), _WEIGHT_EPOCH[0], pair)
    return pair


class _HeadFnFp32(torch.autograd.Function):
    """The projection head in the fp32-faithful mode: every contraction (2 forward, 4 backward) is a three-GEMM
    product of hi/lo splits; GELU and the bias sums stay in fp32."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        xs = ops.split_bf16(x.detach())
        w1s, w2s = split_weight(w1), split_weight(w2)
        a = ops.gemm3(xs, w1s, bias_n=b1.detach())
        hs = ops.split_bf16(ops.gelu_fwd_f32(a))
        z = ops.gemm3(hs, w2s, bias_n=b2.detach())
        ctx.save_for_backward(a, *xs, *hs, *w1s, *w2s)
        ctx.in_dtype = x.dtype
        return z

    @staticmethod
    def backward(ctx, dz):
        a, xh, xl, hh, hl, w1h, w1l, w2h, w2l = ctx.saved_tensors
        dzf = dz if (dz.dtype == torch.float32 and dz.stride(-1) == 1) else dz.float().contiguous()
        dzs = ops.split_bf16(dzf)
        db2 = ops.cols_sum(dzf)
        dw2 = ops.gemm3(dzs, (hh, hl), a_mn_major=True, b_mn_major=True)          # dz^T h   (K, D)
        dh = ops.gemm3(dzs, (w2h, w2l), b_mn_major=True)                          # dz W2    (rows, D)
        da, part = ops.gelu_bwd_f32(dh, a)
        db1 = ops.cols_sum(part)
        das = ops.split_bf16(da)
        dw1 = ops.gemm3(das, (xh, xl), a_mn_major=True, b_mn_major=True)          # da^T x   (D, D)
        dx = ops.gemm3(das, (w1h, w1l), b_mn_major=True)                          # da W1    (rows, D)
        return dx.to(ctx.in_dtype), dw1, db1, dw2, db2


class _GramMatrixFp32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats):
        xn, inv = ops.normalize_tokens_f32(feats, skip=0)
        xs = ops.split_bf16(xn)
        gram = ops.gemm3_batched(xs, xs)
        ctx.save_for_backward(feats, *xs, inv)
        return gram

    @staticmethod
    def backward(ctx, dg):
        feats, xh, xl, inv = ctx.saved_tensors
        Bt, T, D = feats.shape
        dgs = ops.split_bf16(dg.float().contiguous())
        dxn = ops.gemm3_batched(dgs, (xh, xl), b_mn_major=True)                                   # dG   @ Xn
        dxn2 = ops.gemm3_batched(dgs, (xh, xl), a_mn_major=True, b_mn_major=True)                 # dG^T @ Xn
        ops.axpby(dxn, 1.0, dxn2, 1.0, out=dxn)
        grad = torch.empty(Bt, T, D, dtype=torch.float32, device=dg.device)
        ops.normalize_tokens_bwd(feats, dxn, inv, grad, skip=0)
        return grad.to(feats.dtype)


class _GramAnchorFp32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student_feats, teacher_feats):
        Bt, T, D = student_feats.shape
        xs_f, inv_s = ops.normalize_tokens_f32(student_feats, skip=1)
        xt_f, _ = ops.normalize_tokens_f32(teacher_feats, skip=1)
        xs, xt = ops.split_bf16(xs_f), ops.split_bf16(xt_f)
        gs, gt = ops.gemm3_batched(xs, xs), ops.gemm3_batched(xt, xt)
        n = Bt * (T - 1) * (T - 1)
        need_grad = student_feats.requires_grad
        loss, delta = ops.sqdiff(gs, gt, 1.0 / n, want_delta=need_grad)
        if need_grad:
            ctx.save_for_backward(student_feats, *xs, inv_s, delta)
            ctx.n = n
        return loss

    @staticmethod
    def backward(ctx, g):
        student_feats, xh, xl, inv_s, delta = ctx.saved_tensors
        Bt, T, D = student_feats.shape
        # dL/dXn = (dG + dG^T) Xn = (4/n) * Delta @ Xn   (Delta symmetric)
        dxn = ops.gemm3_batched(ops.split_bf16(delta), (xh, xl), b_mn_major=True, alpha=4.0 / ctx.n)
        grad = torch.empty(Bt, T, D, dtype=torch.float32, device=g.device)
        up = g.to(torch.float32).reshape(1).contiguous()
        ops.normalize_tokens_bwd(student_feats, dxn, inv_s, grad, skip=1, scale_dev=up)
        return grad.to(student_feats.dtype), None


class _HeadFn(torch.autograd.Function):
    """logits = W2 . gelu(W1 . x + b1) + b2 with bf16 tensor-core operands / fp32 accumulation, and
    the matching backward GEMMs (operands taken in place through MN-major descriptors)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, out_dtype):
        xb = _to_bf16_rows(x.detach())
        w1b, w2b = bf16_weight(w1), bf16_weight(w2)
        a = ops.gemm_bf16(xb, w1b, bias_n=b1.detach())
        h = ops.gelu_fwd(a)
        z = ops.gemm_bf16(h, w2b, bias_n=b2.detach(), out_dtype=out_dtype)
        ctx.save_for_backward(xb, a, h, w1b, w2b)
        ctx.in_dtype = x.dtype
        return z

    @staticmethod
    def backward(ctx, dz):
        xb, a, h, w1b, w2b = ctx.saved_tensors
        dzb = dz if dz.dtype == torch.bfloat16 and dz.stride(1) == 1 else _to_bf16_rows(dz)
        db2 = ops.cols_sum(dzb)
        dw2 = ops.gemm_bf16(dzb, h, a_mn_major=True, b_mn_major=True)        # dz^T h   (K, D)
        dh = ops.gemm_bf16(dzb, w2b, b_mn_major=True)                         # dz W2    (rows, D)
        da, part = ops.gelu_bwd(dh, a)
        db1 = ops.cols_sum(part)
        dw1 = ops.gemm_bf16(da, xb, a_mn_major=True, b_mn_major=True)         # da^T x   (D, D)
        dx = ops.gemm_bf16(da, w1b, b_mn_major=True)                          # da W1    (rows, D)
        return dx.to(ctx.in_dtype), dw1, db1, dw2, db2, None


class ProjectionHead(nn.Sequential):
    """nn.Sequential(Linear(D,D), GELU(), Linear(D,K)) - identical parameters / state-dict keys
    (`0.weight`, `0.bias`, `2.weight`, `2.bias`) to zoo/arch.py:252-256, forward on tcgen05 GEMMs."""

    def __init__(self, dim: int, out_dim: int) -> None:
        super().__init__(nn.Linear(dim, dim), nn.GELU(), nn.Linear(dim, out_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise _ext.DinoxError("dinox_b200: CUDA tensors required (no CPU fallback)")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if _fp32_mode():
            z = _HeadFnFp32.apply(x2, self[0].weight, self[0].bias, self[2].weight, self[2].bias)
            return z.reshape(*lead, z.shape[-1])
        out_dtype = torch.bfloat16 if torch.is_autocast_enabled() else torch.float32
        z = _HeadFn.apply(x2, self[0].weight, self[0].bias, self[2].weight, self[2].bias, out_dtype)
        return z.reshape(*lead, z.shape[-1])


class DinoStudentTeacher(nn.Module):
    """DINO student/teacher wrapper with projection head (zoo/arch.py:246-261)."""

    def __init__(self, backbone: nn.Module, out_dim: int = 8192) -> None:
        super().__init__()
        self.backbone = backbone
        self.head = ProjectionHead(backbone.dim, out_dim)

    def forward(self, x: torch.Tensor, spacing: Optional[torch.Tensor] = None) -> torch.Tensor:
        feats = self.backbone(x, spacing=spacing)
        return self.head(feats[:, 0])


# =================================================================================================
# Fused path: head + multi-crop CE + iBOT with logits living only in TMEM
# =================================================================================================
class _EntryPlan:
    """Host-built (cached) index tables pairing student rows with teacher rows.

    CLS entries: every (teacher view iq, student view v != iq, image b) -> student row v*B+b,
    teacher row iq*B+b, weight 1/(n_terms*B); padded to a multiple of 128.  iBOT entries: masked
    token m -> student row Ms+m, teacher row Mt+m (weights come from masks_weight on device)."""

    def __init__(self, B: int, Vg: int, V: int, Mm: int, device):
        Ms, Mt = B * V, B * Vg
        es, et = [], []
        for iq in range(Vg):
            for v in range(V):
                if v == iq:
                    continue
                for b in range(B):
                    es.append(v * B + b)
                    et.append(iq * B + b)
        self.n_cls = len(es)
        self.n_terms = Vg * V - Vg
        pad = (-self.n_cls) % 128
        es += [-1] * pad
        et += [-1] * pad
        self.e_cls_pad = len(es)
        es += [Ms + m for m in range(Mm)]
        et += [Mt + m for m in range(Mm)]
        self.E = len(es)
        self.e_pad = (self.E + 127) // 128 * 128
        pad2 = self.e_pad - self.E
        es += [-1] * pad2
        et += [-1] * pad2
        # CSR: entries of every student row (for dH_row = sum of its entries' dH)
        rows = Ms + Mm
        buckets: List[List[int]] = [[] for _ in range(rows)]
        for e, r in enumerate(es):
            if r >= 0:
                buckets[r].append(e)
        ptr = [0]
        ent = []
        for bkt in buckets:
            ent += bkt
            ptr.append(len(ent))
        cw = [1.0 / (self.n_terms * B)] * self.n_cls + [0.0] * (self.e_pad - self.n_cls)
        t = lambda x, dt: torch.tensor(x, dtype=dt, device=device)
        self.ent_s, self.ent_t = t(es, torch.int64), t(et, torch.int64)
        self.csr_ptr, self.csr_ent = t(ptr, torch.int64), t(ent, torch.int64)
        self.cw_base = t(cw, torch.float32)
        self.Ms, self.Mt, self.Mm, self.B, self.V, self.Vg = Ms, Mt, Mm, B, V, Vg
        # read-back path: the teacher rows are laid out [CLS rows | zero rows up to a multiple of 128 | masked
        # patch rows] so that one teacher launch can switch its column offsets (centre / patch centre) per M tile
        self.Mt_pad = (Mt + 127) // 128 * 128 if Mm else Mt
        etp = [(r if r < Mt else r - Mt + self.Mt_pad) if r >= 0 else -1 for r in et]
        self.ent_t_pad = t(etp, torch.int64)                           # -1 = padding entry
        self.trow = t([max(r, 0) for r in etp], torch.int32)           # row of qt / refs (padding -> row 0, weight 0)
        self.srow = t(es, torch.int32)                                 # student row of the entry (-1 = padding)
        self.cls_rows_pad = t(list(range(Mt)) + [-1] * (self.Mt_pad - Mt), torch.int64)
        self.row_counts = t([float(Mt), float(Mm), 0.0, 0.0], torch.float32)   # teacher CLS rows, masked patch rows


_PLANS: Dict[Tuple, _EntryPlan] = {}


def _entry_plan(B, Vg, V, Mm, device) -> _EntryPlan:
    key = (B, Vg, V, Mm, str(device))
    p = _PLANS.get(key)
    if p is None:
        p = _EntryPlan(B, Vg, V, Mm, device)
        _PLANS[key] = p
    return p


_SIDE_STREAMS: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}


def concurrency() -> int:
    """DINOX_CONCURRENCY: 0 = one stream; 1 = Gram anchoring on a side stream (LossHeadStep);
    2 (default) = also the teacher branch of the fused head loss; 3 = also dW2/db2 beside the dH chain in the
    backward and the centre GEMV inside the teacher branch."""
    return int(os.environ.get("DINOX_CONCURRENCY", "2"))


def stream_priorities() -> bool:
    """DINOX_STREAM_PRIO (default 1): the streams of the critical chain (teacher branch, backward side chains, the
    capture stream of LossHeadStep) get a higher CUDA priority than the Gram-anchoring and centre-update streams, so
    that pending blocks of the chain are dispatched first when SMs free up."""
    return os.environ.get("DINOX_STREAM_PRIO", "1") != "0"


# which: 0 teacher branch, 1 Gram anchoring, 2 dW2/db2 at concurrency level 3, 3 centre update, 4 / 6 parameter-gradient
# chains of the backward, 5 entry staging beside pass 1
_LOW_PRIORITY_STREAMS = (1, 3)


def _side_stream(device, which: int = 0) -> "torch.cuda.Stream":
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), which)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        prio = -1 if (stream_priorities() and which not in _LOW_PRIORITY_STREAMS) else 0
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev, priority=prio)
    return st


def pass2_mode() -> str:
    """DINOX_PASS2: "readback" (default) - the teacher runs ONE pass that also writes its un-normalised fp16
    probabilities (dinox_head_teacher) and pass 2 recomputes only the student logits (dinox_head_grad2);
    "recompute" - round-1 path: teacher statistics pass + both logit tiles recomputed side by side in pass 2."""
    m = os.environ.get("DINOX_PASS2", "readback")
    if m not in ("readback", "recompute"):
        raise ValueError(f"DINOX_PASS2={m!r}: expected readback or recompute")
    return m


def _accumulate_grad(p: torch.Tensor, fn) -> None:
    """fn(out_tensor, accumulate: bool) writes/accumulates dL/dp straight into p.grad (fp32)."""
    if p.grad is None:
        if torch.cuda.is_current_stream_capturing():
            raise _ext.DinoxError("graph capture needs pre-allocated .grad tensors on the head parameters "
                                  "(LossHeadStep allocates and zeroes them)")
        p.grad = torch.empty_like(p, memory_format=torch.contiguous_format)
        fn(p.grad, False)
    else:
        fn(p.grad, True)


def _update_centres(stats, work, w2t, b2t, loss_mod, center_patch, patch_momentum, cls: bool, patch: bool) -> None:
    """Centre EMA of the fused path: mean teacher logits = W2t . mean(h_t) + b2t (one GEMV over W2t for the CLS
    and the patch centre together).  `stats` = (hsum (n, D), counts (n,)): the all-reduced activation sums and row
    counts of [CLS rows (if cls)] [masked patch rows (if patch)] - the division by the GLOBAL count happens on the
    device, so ranks may mask different numbers of tokens."""
    if stats is None:
        return
    hsum, counts = stats
    if isinstance(work, tuple):       # ("late", buffer): the all-reduce was deferred to this point
        allreduce_sum_(work[1], loss_mod.process_group)
    elif work is not None:
        work.wait()
    targets = ([loss_mod.center.view(-1)] if cls else []) + ([center_patch.view(-1)] if patch else [])
    momenta = ([loss_mod.center_momentum] if cls else []) + ([patch_momentum] if patch else [])
    # mean logits and both centre EMAs in ONE pass over W2t (scripts/phase5_big_run.py:686-690)
    ops.gemv_bf16_multi_ema_(w2t, hsum, [1.0] * hsum.shape[0], b2t, targets, momenta, divisors=counts)


class _FusedHeadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student_cls, student_patch, w1, b1, w2, b2, teacher_cls, teacher_patch, masks_weight, t_head,
                loss_mod, center_patch, cfg, patch_index=None, params_in_place=None, teacher_index=None):
        (student_temp, teacher_temp, Vg, n_local, ibot_weight, teacher_mode, sk_iters, pg, update_center,
         patch_momentum, grads_in_place, w2_sink, patch_mode) = cfg
        dev = student_cls.device
        D = student_cls.shape[1]
        K = w2.shape[0]
        Mt = teacher_cls.shape[0]
        B = Mt // Vg
        V = student_cls.shape[0] // B
        Ms = B * V
        # iBOT rows: materialised (Mm, D) rows, or - with patch_index - gathered in the staging kernel straight
        # from the backbone's token tensors (any (..., D) contiguous shape; patch_index = flat row numbers)
        Mm = 0 if student_patch is None else (student_patch.shape[0] if patch_index is None else patch_index.numel())
        sp2 = tp2 = None
        if Mm:
            sp2 = student_patch.detach() if patch_index is None else student_patch.detach().view(-1, D)
            if teacher_index is None:
                teacher_index = patch_index
            tp2 = teacher_patch.detach() if teacher_index is None else teacher_patch.detach().view(-1, D)
        plan = _entry_plan(B, Vg, V, Mm, dev)
        readback = pass2_mode() == "readback"
        Mt_pad = plan.Mt_pad if readback else Mt        # row of the first masked patch in the teacher matrices
        inv_ts, inv_tt = 1.0 / student_temp, 1.0 / teacher_temp
        # the bf16 copies are cached per PARAMETER object: in in-place mode w1 / w2 are fresh detached views on every
        # call (a cache miss = a 150 MB cast of W2 inside every step), the real parameters ride in params_in_place
        pw1, pw2 = (params_in_place[0], params_in_place[2]) if params_in_place is not None else (w1, w2)
        w1s, w2s = bf16_weight(pw1), bf16_weight(pw2)
        w1t, w2t = bf16_weight(t_head[0].weight), bf16_weight(t_head[2].weight)
        centre_cls = update_center and teacher_mode == "center"
        # the patch centre is the only collapse protection of the iBOT targets in BOTH teacher modes (the CLS
        # Sinkhorn-Knopp normalisation does not touch the patch rows), so it is updated whenever there are patch rows
        centre_patch = update_center and Mm > 0 and patch_mode == "center"

        # Two branches that meet at pass 2.  The teacher branch (no gradient) is issued on a side stream
        # so that its short kernels slot into the launch gaps and wave tails of the student branch;
        # per-kernel timing (ops.TIMER) keeps everything on one stream.
        main = torch.cuda.current_stream()
        side = _side_stream(dev) if (concurrency() >= 2 and not ops.TIMER.enabled) else None
        # per-prototype offsets of both branches in one launch, before the streams fork (the centre teacher of the CLS
        # rows and the centred patch teacher; Sinkhorn modes build theirs from the iteration results below)
        pre = None
        if teacher_mode == "center" and (not Mm or patch_mode == "center"):
            pre = ops.head_offsets(b2.detach(), t_head[2].bias.detach(), loss_mod.center.reshape(-1),
                                   center_patch.reshape(-1) if Mm else None, inv_ts, inv_tt)
        if side is not None:
            side.wait_stream(main)
        qt = refs = ht_e = None
        # ---- teacher: stage inputs [CLS rows | (zero rows) | masked patch rows] as bf16, layer 1, statistics
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            xt = torch.empty(Mt_pad + Mm, D, dtype=torch.bfloat16, device=dev)
            cls_idx = plan.cls_rows_pad if Mt_pad != Mt else None
            if Mm:   # [CLS rows | zero rows | masked patch rows] staged by one launch
                ops.gather_cast_bf16_2(teacher_cls.detach(), cls_idx, Mt_pad, tp2, teacher_index, xt)
            else:
                ops.gather_cast_bf16(teacher_cls.detach(), cls_idx, xt)
            a_t = ops.gemm_bf16(xt, w1t, bias_n=t_head[0].bias.detach())   # layer 1 (zoo/arch.py:253-254)
            ht = ops.gelu_fwd(a_t)
            del a_t
            # centre statistics: batch-mean teacher logits are W2t . mean(h_t) + b2t by
            # linearity, so the data-parallel payload is the D-vector sum(h_t) (CLS and masked-patch rows
            # in ONE all-reduce), launched now so that its latency hides behind the pass-1/pass-2 GEMMs
            stats = hsum_work = None
            if centre_cls or centre_patch:
                nv = int(centre_cls) + int(centre_patch)
                # [nv x D activation sums | nv row counts] in one buffer = ONE all-reduce; the counts make the mean
                # exact when ranks hold different numbers of rows (masked tokens)
                sbuf = torch.empty(nv * D + 4, dtype=torch.float32, device=dev)
                hsum, counts = sbuf[:nv * D].view(nv, D), sbuf[nv * D:nv * D + nv]
                i0 = 0 if centre_cls else 1
                segs = ([(0, Mt)] if centre_cls else []) + ([(Mt_pad, Mt_pad + Mm)] if centre_patch else [])
                # activation sums of both row ranges and their row counts in one launch
                ops.segment_cols_sum(ht, segs, hsum, plan.row_counts[i0:i0 + nv], counts)
                # DINOX_CENTER_AR=late issues the all-reduce only where its result is needed (after pass 2) instead of
                # here, next to the teacher pass (A/B knob: does the NCCL kernel disturb the persistent GEMMs?)
                if os.environ.get("DINOX_CENTER_AR", "early") == "late":
                    hsum_work = ("late", sbuf)
                else:
                    hsum_work = allreduce_sum_async(sbuf, pg)
                stats = (hsum, counts)
            b2t = t_head[2].bias.detach()
            center = loss_mod.center.reshape(-1)
            rb2_t = torch.empty(Mt_pad + Mm, dtype=torch.float32, device=dev)
            b_row = None
            if pre is not None:
                ct2 = pre[1]
            elif teacher_mode == "center":
                ct2 = ops.axpby(b2t, inv_tt * LOG2E, center, -inv_tt * LOG2E)
            else:  # Sinkhorn-Knopp on the (small) materialised CLS teacher logits
                t_cls = ops.gemm_bf16(ht[:Mt], w2t, bias_n=b2t)
                a_col, b_row = sinkhorn_knopp_biases(t_cls, teacher_temp, sk_iters, pg)
                ct2 = ops.axpby(b2t, inv_tt * LOG2E, a_col, -LOG2E)
                del t_cls
            b_row_patch = None
            if Mm and patch_mode == "sinkhorn":
                # Sinkhorn-Knopp over the masked patches (DINOv2): the iterations need the whole (Mm, K) logit matrix
                # three times over, so it is materialised once (fp32) for them; the resulting per-prototype offsets go
                # into the teacher pass as its column offsets and the per-token offsets replace the row LSE
                t_patch = ops.gemm_bf16(ht[Mt_pad:], w2t, bias_n=b2t)
                a_col_p, b_row_patch = sinkhorn_knopp_biases(t_patch, teacher_temp, sk_iters, pg)
                ct2_patch = ops.axpby(b2t, inv_tt * LOG2E, a_col_p, -LOG2E)
                del t_patch
            elif pre is not None:
                ct2_patch = pre[2]
            else:
                ct2_patch = ops.axpby(b2t, inv_tt * LOG2E, center_patch.reshape(-1), -inv_tt * LOG2E) if Mm else None
            # The centre EMA (a 50 MB GEMV over W2t) only needs the activation sums: from here on - the offsets above hold
            # copies of the OLD centres, scripts/phase5_big_run.py:719 - it runs on its own stream beside the three big
            # passes and is joined after pass 2 (it used to sit between pass 2 and the backward: 36 us of the step)
            centre_stream = None
            if side is not None and stats is not None and concurrency() < 3:
                centre_stream = _side_stream(dev, 3)
                centre_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(centre_stream):
                    _update_centres(stats, hsum_work, w2t, b2t, loss_mod, center_patch, patch_momentum, centre_cls, centre_patch)
                stats = None
            if readback:
                # ONE pass over the teacher rows: statistics + un-normalised fp16 probabilities
                with ops.TIMER.region("head_teacher"):
                    qt, refs, _ = ops.head_teacher(ht, w2t, inv_tt, ct2, ct2_patch, Mt_pad, out_log2=rb2_t)
            else:
                if teacher_mode == "center":
                    with ops.TIMER.region("head_stats_teacher_cls"):
                        ops.head_stats(ht[:Mt], w2t, inv_tt, ct2, want_nat=False, out_log2=rb2_t[:Mt])
                if Mm:
                    with ops.TIMER.region("head_stats_teacher_patch"):
                        ops.head_stats(ht[Mt:], w2t, inv_tt, ct2_patch, want_nat=False, out_log2=rb2_t[Mt:])
                ht_e = torch.empty(plan.e_pad, D, dtype=torch.bfloat16, device=dev)
                ops.gather_cast_bf16(ht, plan.ent_t, ht_e)
            if b_row is not None:   # Sinkhorn rows are normalised by their own offset, not by the row LSE
                ops.axpb(b_row, LOG2E, out=rb2_t[:Mt])
            if b_row_patch is not None:
                ops.axpb(b_row_patch, LOG2E, out=rb2_t[Mt_pad:])
            # read-back pass 2 looks the row statistics up itself (entry -> row tables of the plan; entries of weight 0
            # are dead).  Round-1 pass 2 takes them per entry: a huge offset makes the padding entries'
            # probabilities exactly 0 (0 * inf would poison the sums)
            rb2_e = rb2_t if readback else ops.gather_f32(rb2_t, plan.ent_t, fill=1.0e30)
            if side is not None and concurrency() >= 3:
                # centre updates here instead of after pass 2: ct2 / ct2_patch above already hold the OLD centres
                # (scripts/phase5_big_run.py:719 - the loss sees the centre of the previous step), and the GEMV over
                # W2t runs beside the student branch instead of after pass 2
                _update_centres(stats, hsum_work, w2t, b2t, loss_mod, center_patch, patch_momentum, centre_cls, centre_patch)
                stats = None
        # ---- student: the same on the current stream
        xs = torch.empty(Ms + Mm, D, dtype=torch.bfloat16, device=dev)
        if Mm:
            ops.gather_cast_bf16_2(student_cls.detach(), None, Ms, sp2, patch_index, xs)
        else:
            ops.gather_cast_bf16(student_cls.detach(), None, xs)
        a_s = ops.gemm_bf16(xs, w1s, bias_n=b1.detach())
        hs = ops.gelu_fwd(a_s)
        b2s = b2.detach()
        cs2 = pre[0] if pre is not None else ops.axpb(b2s, inv_ts * LOG2E)
        # ---- entries: the student activations per entry and the entry weights only need the GELU output - they are
        # staged beside pass 1 (own stream; these small kernels fit on the SMs next to its CTAs) instead of between the
        # teacher pass and pass 2
        aux = _side_stream(dev, 5) if side is not None else None
        if aux is not None:
            aux.wait_stream(main)
        with (torch.cuda.stream(aux) if aux is not None else contextlib.nullcontext()):
            hs_e = torch.empty(plan.e_pad, D, dtype=torch.bfloat16, device=dev)
            ops.gather_cast_bf16(hs, plan.ent_s, hs_e)
            cw = plan.cw_base
            if Mm:
                mw = masks_weight.detach()
                mw = mw if (mw.dtype == torch.float32 and mw.is_contiguous()) else mw.float().contiguous()
                cw = ops.entry_weights(plan.cw_base, mw, plan.e_cls_pad, ibot_weight / Mt)
        with ops.TIMER.region("head_stats_student"):
            _, lse2_s = ops.head_stats(hs, w2s, inv_ts, cs2, want_nat=False)
        lse2_e = lse2_s if readback else ops.gather_f32(lse2_s, plan.ent_s, fill=1.0e30)
        if side is not None:
            # Everything the side branches allocated lives in their streams' pools and is read by pass 2 on
            # this stream: those blocks are only handed out again to side-stream work, and every side-stream
            # region starts by waiting for this stream, i.e. after pass 2 has been queued.
            main.wait_stream(side)
            main.wait_stream(aux)
        # ---- pass 2
        lbuf = torch.empty(4, dtype=torch.float32, device=dev)     # [L_dino, L_ibot, total, -]
        losses = lbuf[:3]
        need_grad = any(ctx.needs_input_grad[:6]) or (params_in_place is not None and
                                                      any(p.requires_grad for p in params_in_place))
        with ops.TIMER.region("head_grad"):
            if readback:
                gt, db2p = ops.head_grad2(hs_e, w2s, inv_ts, cs2, lse2_e, cw, rb2_e, plan.trow, qt, refs, plan.e_cls_pad,
                                          losses, want_db2=need_grad, srow_e=plan.srow)
            else:
                gt, db2p = ops.head_grad(w2s, w2t, hs_e, ht_e, inv_ts, inv_tt, cs2, ct2, ct2_patch, plan.e_cls_pad,
                                         lse2_e, rb2_e, cw, losses, want_db2=need_grad)
        # ---- centre updates AFTER the loss (scripts/phase5_big_run.py:719)
        _update_centres(stats, hsum_work, w2t, b2t, loss_mod, center_patch, patch_momentum, centre_cls, centre_patch)
        if centre_stream is not None:
            main.wait_stream(centre_stream)
        if need_grad:
            ctx.save_for_backward(xs, a_s, hs_e, gt, db2p, w1s, w2s)
            ctx.plan = plan
            ctx.params = params_in_place if params_in_place is not None else (w1, b1, w2, b2)
            ctx.readback, ctx.grads_in_place, ctx.w2_sink = readback, grads_in_place, w2_sink
            ctx.in_dtypes = (student_cls.dtype, None if student_patch is None else student_patch.dtype)
            ctx.patch_index = patch_index
            ctx.patch_shape = None if student_patch is None else tuple(student_patch.shape)
        if readback:
            total = lbuf[2]                      # written by pass 2 itself
        else:
            total = ops.scalar_combine([lbuf[0], lbuf[1]], [1.0, 1.0])
        losses = lbuf[:2]
        ctx.mark_non_differentiable(losses)
        ctx.set_materialize_grads(False)     # no zero-filled gradient for the (non-differentiable) loss pair
        return total, losses

    @staticmethod
    def backward(ctx, g, _g_losses):
        if g is None:
            return (None,) * 16
        xs, a_s, hs_e, gt, db2p, w1s, w2s = ctx.saved_tensors
        plan = ctx.plan
        w1, b1, w2, b2 = ctx.params
        in_place = ctx.grads_in_place
        up = g.to(torch.float32).reshape(1).contiguous()
        K, D = w2s.shape
        rows = plan.Ms + plan.Mm
        dev = gt.device
        outs = {}

        def emit(name, p, fn):
            """dL/dp: accumulated straight into p.grad (in-place mode) or returned to autograd"""
            if in_place:
                _accumulate_grad(p, fn)
            else:
                o = torch.empty_like(p, memory_format=torch.contiguous_format)
                fn(o, False)
                outs[name] = o

        # The layer-2 parameter gradients (dW2, db2) and the dH -> layer-1 chain only share their input G: at
        # level 3 the former run on a side stream, so the CTAs of dH start on the SMs the last, partial wave of
        # dW2 leaves idle (256 tile pairs on 74 CTA pairs = 3.46 waves).
        main = torch.cuda.current_stream()
        side = _side_stream(dev, 2) if (concurrency() >= 3 and not ops.TIMER.enabled) else None
        if side is not None:
            side.wait_stream(main)
        rb = ctx.readback   # G is (E, K) entry-major on the read-back path, Gt (K, E) prototype-major otherwise
        # DINOX_BALANCED (bit mask, default 0): 1 = dW2, 2 = dH on the balanced schedule (dinox_gemm_bf16_balanced).
        # Measured at C2: dW2 alone 0.335 -> 0.318 ms, dH 0.296 -> 0.308 ms, but the micro-step does not get shorter
        # (2.322 vs 2.337 ms, 3 interleaved runs each): inside the step the idle SMs of dW2's last wave are already
        # taken by the side-stream kernels (centre GEMV, Gram backward) - profiles/README.md.
        balanced = int(os.environ.get("DINOX_BALANCED", "0"))
        # the bias gradients and dW1 never feed another kernel of this backward: they run on a side stream beside the
        # dW2 -> dH -> dL/dx chain (db2's 70 MB column sum used to stand between dW2 and dH)
        side1 = _side_stream(dev, 4) if (concurrency() >= 2 and not ops.TIMER.enabled) else None
        if side1 is not None:
            side1.wait_stream(main)
            with torch.cuda.stream(side1):
                emit("b2", b2, lambda out, acc: ops.cols_sum_axpy_(db2p, out, acc, scale_dev=up))
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            # dW2 (K, D) += g * G^T . HsE   (A = G with the prototypes as M; B = HsE MN-major)
            with ops.TIMER.region("gemm_dW2"):
                sink = ctx.w2_sink
                if sink is not None and sink.flush:
                    # data parallel, LAST micro-step of the accumulation window: every output tile - plus the gradient
                    # accumulated locally over the earlier micro-steps - is reduce-added into the OWNER rank's
                    # peer-mapped shard (mean over ranks): GEMM and reduce-scatter in one kernel, one gradient per
                    # window over NVLink
                    local = w2.grad
                    ops.gemm_bf16_reduce_scatter(gt, hs_e, sink.ptrs, sink.rows, sink.cols, a_mn_major=rb, b_mn_major=True,
                                                 alpha=1.0 / sink.world, alpha_dev=up, add_local=local,
                                                 add_scale=1.0 / sink.world)
                    if local is not None:
                        w2.grad = None    # shipped: the owner's shard holds it now
                elif balanced & 1:
                    # 256 tile pairs on 74 CTA pairs: the 34 tiles of the last wave are cut in two along the entries
                    emit("w2", w2, lambda out, acc: ops.gemm_bf16_balanced(
                        gt, hs_e, tag="dW2", a_mn_major=rb, b_mn_major=True, out=out, accumulate=acc, alpha_dev=up,
                        m_fastest=False))
                else:
                    emit("w2", w2, lambda out, acc: ops.gemm_bf16(
                        gt, hs_e, a_mn_major=rb, b_mn_major=True, out=out, accumulate=acc, alpha_dev=up, m_fastest=False))
            if side1 is None:
                emit("b2", b2, lambda out, acc: ops.cols_sum_axpy_(db2p, out, acc, scale_dev=up))
        # dH per entry = G . W2  (B = W2 MN-major), then sum the entries of each row
        with ops.TIMER.region("gemm_dH"):
            if balanced & 2:   # fewer tiles than CTA pairs: every tile cut along the prototypes, accumulated in place
                dh_e = ops.gemm_bf16_balanced(gt, w2s, tag="dH", a_mn_major=not rb, b_mn_major=True, m_fastest=True)
            else:
                dh_e = ops.gemm_bf16_splitk(gt, w2s, a_mn_major=not rb, b_mn_major=True, m_fastest=True)
        # dL/dh of a row = sum of its entries' rows (and of the split-K slabs), formed inside the GELU backward
        da, part = ops.gelu_bwd_gather(dh_e, plan.csr_ptr, plan.csr_ent, a_s, scale_dev=up)
        # the layer-1 parameter gradients (db1, dW1) and dL/dx only share da: two short chains side by side
        if side1 is not None:
            side1.wait_stream(main)
        side2 = _side_stream(dev, 6) if side1 is not None else None
        if side2 is not None:
            side2.wait_stream(main)
        with (torch.cuda.stream(side2) if side2 is not None else contextlib.nullcontext()):
            emit("b1", b1, lambda out, acc: ops.cols_sum_axpy_(part, out, acc))
        with (torch.cuda.stream(side1) if side1 is not None else contextlib.nullcontext()):
            # dW1 = da^T x: 3x3 output tiles with a reduction over every row -> split-K, slabs summed in fixed order
            dw1_parts = ops.gemm_bf16_splitk(da, xs, a_mn_major=True, b_mn_major=True)
            emit("w1", w1, lambda out, acc: ops.sum_slabs(dw1_parts, out, accumulate=acc))
        dx = ops.gemm_bf16(da, w1s, b_mn_major=True)
        if side1 is not None:
            main.wait_stream(side1)
            main.wait_stream(side2)
        if side is not None:
            main.wait_stream(side)
        d_cls = dx[:plan.Ms].to(ctx.in_dtypes[0]) if ctx.needs_input_grad[0] else None
        d_patch = None
        if plan.Mm and ctx.needs_input_grad[1]:
            if ctx.patch_index is None:
                d_patch = dx[plan.Ms:].to(ctx.in_dtypes[1])
            else:   # rows went in through an index: their gradients go back to those rows of the token tensor
                d_tok = torch.empty(ctx.patch_shape, dtype=torch.float32, device=dx.device)
                ops.fill_(d_tok.view(-1), 0.0)
                ops.scatter_rows(dx[plan.Ms:], ctx.patch_index, d_tok.view(-1, D))
                d_patch = d_tok.to(ctx.in_dtypes[1])
        return (d_cls, d_patch, outs.get("w1"), outs.get("b1"), outs.get("w2"), outs.get("b2"),
                None, None, None, None, None, None, None, None, None, None)


def fused_head_dino_loss(student_cls: torch.Tensor, teacher_cls: torch.Tensor, student_head: nn.Sequential,
                         teacher_head: nn.Sequential, dino_loss: DINOLoss, student_temp: float, teacher_temp: float,
                         *, student_patch: Optional[torch.Tensor] = None, teacher_patch: Optional[torch.Tensor] = None,
                         masks_weight: Optional[torch.Tensor] = None, center_patch: Optional[torch.Tensor] = None,
                         ibot_weight: float = 1.0, patch_center_momentum: Optional[float] = None,
                         update_center: bool = True, patch_index: Optional[torch.Tensor] = None,
                         grads_in_place: bool = False,
                         teacher_patch_index: Optional[torch.Tensor] = None,
                         w2_grad_shards=None) -> Dict[str, torch.Tensor]:
    """Projection head + multi-crop DINO CE (+ iBOT masked-patch CE) in one fused path.

    Equivalent to `dino_loss(student_head(student_cls), teacher_head(teacher_cls), ...)` of the reference
    loop (scripts/phase5_big_run.py:1746-1754) but the (rows x K) logits exist only in TMEM.
    Gradients flow through autograd to `student_cls` / `student_patch` AND to the four student head
    parameters (AccumulateGrad hooks, DDP reducers and `torch.autograd.grad` all see them).
    `grads_in_place=True` is the opt-in fast path of a single-process accumulation loop: the head gradients
    are reduce-added straight into the parameters' `.grad` by the GEMM epilogues (no (K, D) temporary, no
    AccumulateGrad pass; scaled by the upstream gradient, so `loss / accumulation_steps` and GradScaler work),
    invisible to autograd hooks.
    With `patch_index` (int64, unique flat row numbers) `student_patch` / `teacher_patch` are the backbone's
    token tensors themselves (contiguous (..., D), e.g. the (B*Vg, T, D) output that `feats[:, 1:]` slices,
    scripts/phase5_big_run.py:1741-1747): the masked rows are gathered by the staging kernel and their
    gradients scattered back, so the caller never materialises `tokens[mask]`.
    `w2_grad_shards` (an `optim.PeerGradShards` of `student_head[2].weight`, data parallel only): dW2 is not
    written to `.grad` / returned; the backward GEMM reduce-adds its tiles into the owner ranks' peer-mapped shards
    (fused GEMM + reduce-scatter over NVLink), to be consumed by `ShardedFusedAdamW(grad_shards=...)`.
    `teacher_patch_index` names the rows of the TEACHER token tensor only (`student_patch` then holds materialised
    (Mm, D) rows): what `token_fork` + `LossHeadStep` use so that the iBOT gradient is added into the Gram-anchoring
    gradient of the same token tensor instead of travelling as a second dense tensor.
    Returns {"loss": differentiable total, "loss_dino", "loss_ibot"}."""
    for t in (student_cls, teacher_cls):
        if not t.is_cuda:
            raise _ext.DinoxError("dinox_b200: CUDA tensors required (no CPU fallback)")
    has_ibot = student_patch is not None
    if has_ibot:
        if teacher_patch is None or masks_weight is None or center_patch is None:
            raise ValueError("iBOT needs teacher_patch, masks_weight and center_patch")
        if patch_index is not None:
            if patch_index.dtype != torch.int64 or patch_index.dim() != 1 or not patch_index.is_cuda:
                raise ValueError("patch_index: 1-D int64 CUDA tensor of flat token-row numbers expected")
            if not (student_patch.is_contiguous() and teacher_patch.is_contiguous()):
                raise ValueError("patch_index needs contiguous token tensors (pass the backbone output, not a slice)")
            if masks_weight.numel() != patch_index.numel():
                raise ValueError("masks_weight and patch_index must name the same rows")
        if teacher_patch_index is not None:
            if patch_index is not None:
                raise ValueError("give either patch_index (both token tensors) or teacher_patch_index (teacher only)")
            ti = teacher_patch_index
            if ti.dtype != torch.int64 or ti.dim() != 1 or not ti.is_cuda or ti.numel() != student_patch.shape[0]:
                raise ValueError("teacher_patch_index: 1-D int64 CUDA tensor with one entry per student_patch row expected")
            if not teacher_patch.is_contiguous():
                raise ValueError("teacher_patch_index needs a contiguous teacher token tensor")
    cfg = (student_temp, teacher_temp, dino_loss.n_global, dino_loss.n_local, ibot_weight, dino_loss.teacher_mode,
           dino_loss.sk_iterations, dino_loss.process_group, update_center,
           dino_loss.center_momentum if patch_center_momentum is None else patch_center_momentum, bool(grads_in_place),
           w2_grad_shards, dino_loss.patch_teacher_mode)
    params = (student_head[0].weight, student_head[0].bias, student_head[2].weight, student_head[2].bias)
    # in-place mode: the parameters enter detached (no AccumulateGrad node takes part in the backward - theirs would
    # tie a captured backward to whatever stream first created them) and the real ones ride along to receive .grad
    ins = tuple(p.detach() for p in params) if grads_in_place else params
    total, losses = _FusedHeadLoss.apply(student_cls, student_patch, *ins,
                                         teacher_cls.detach(), None if teacher_patch is None else teacher_patch.detach(),
                                         masks_weight, teacher_head, dino_loss, center_patch, cfg, patch_index,
                                         params if grads_in_place else None, teacher_patch_index)
    return {"loss": total, "loss_dino": losses[0], "loss_ibot": losses[1]}


class _TokenFork(torch.autograd.Function):
    """One token tensor feeding two loss terms: returns (the tensor itself for Gram anchoring, the masked rows as a
    materialised (Mm, D) fp32 matrix for the iBOT term).  Backward: the dense gradient that arrives for the first
    output (Gram anchoring owns it) receives the row gradients of the second by an in-place scatter-add - no
    zero-filled dense tensor for the rows, no framework add of two dense tensors."""

    @staticmethod
    def forward(ctx, tokens, index):
        D = tokens.shape[-1]
        flat = tokens.detach().reshape(-1, D)
        rows = torch.empty(index.numel(), D, dtype=torch.float32, device=tokens.device)
        ops.gather_rows_f32(flat, index, rows)
        ctx.save_for_backward(index)
        ctx.shape, ctx.dtype = tuple(tokens.shape), tokens.dtype
        ctx.set_materialize_grads(False)
        return tokens.view_as(tokens), rows

    @staticmethod
    def backward(ctx, g_tokens, g_rows):
        (index,) = ctx.saved_tensors
        D = ctx.shape[-1]
        if g_tokens is None and g_rows is None:
            return None, None
        if g_tokens is None:
            g_tokens = torch.empty(ctx.shape, dtype=torch.float32, device=index.device)
            ops.fill_(g_tokens.view(-1), 0.0)
        elif not (g_tokens.dtype == torch.float32 and g_tokens.is_contiguous()):
            g_tokens = g_tokens.float().contiguous()
        if g_rows is not None:
            ops.scatter_add_rows(g_rows.float().contiguous(), index, g_tokens.view(-1, D))
        return g_tokens.to(ctx.dtype), None


def token_fork(tokens: torch.Tensor, index: torch.Tensor):
    """(tokens, tokens.reshape(-1, D)[index] as fp32 rows) with a backward that adds the row gradients into the
    dense gradient of the first output in place.  `index`: unique flat row numbers (int64, CUDA)."""
    if not (tokens.is_cuda and tokens.is_contiguous()):
        raise ValueError("token_fork: contiguous CUDA token tensor expected")
    return _TokenFork.apply(tokens, index)


class FusedLossHead(nn.Module):
    """The fused path as ONE module, so that wrappers which hook `forward` see it - in particular
    `torch.nn.parallel.DistributedDataParallel(FusedLossHead(...))`: DDP arms its gradient reducer in `forward`,
    and the head gradients reach it through autograd (default `grads_in_place=False`), i.e. the head's dW1/db1/
    dW2/db2 are all-reduced like every other parameter (gradient all-reduce is the host's DDP job).

    Holds the student head (trainable), the teacher head (frozen; update it with `_ema_update`), the DINOLoss
    state (`center`) and the iBOT patch centre.  `forward` returns the differentiable total (tensor) - DDP wants
    a tensor output - and keeps the component losses in `last`."""

    def __init__(self, student_head: nn.Sequential, teacher_head: nn.Sequential, dino_loss: DINOLoss,
                 ibot_weight: float = 1.0, grads_in_place: bool = False) -> None:
        super().__init__()
        self.student_head, self.teacher_head, self.dino_loss = student_head, teacher_head, dino_loss
        for p in self.teacher_head.parameters():
            p.requires_grad_(False)
        self.register_buffer("center_patch", torch.zeros_like(dino_loss.center))
        self.ibot_weight, self.grads_in_place = ibot_weight, grads_in_place
        self.last: Dict[str, torch.Tensor] = {}

    def forward(self, student_cls, teacher_cls, student_temp: float, teacher_temp: float, student_patch=None,
                teacher_patch=None, masks_weight=None, patch_index=None) -> torch.Tensor:
        out = fused_head_dino_loss(student_cls, teacher_cls, self.student_head, self.teacher_head, self.dino_loss,
                                   student_temp, teacher_temp, student_patch=student_patch, teacher_patch=teacher_patch,
                                   masks_weight=masks_weight,
                                   center_patch=self.center_patch if student_patch is not None else None,
                                   ibot_weight=self.ibot_weight, patch_index=patch_index,
                                   grads_in_place=self.grads_in_place)
        self.last = {k: v.detach() for k, v in out.items()}
        return out["loss"]


class _CombineLosses(torch.autograd.Function):
    """loss = scale * sum_i w_i * term_i in one launch (and one for the backward fan-out) instead of the
    mul/add/div chain of scripts/phase5_big_run.py:1749-1772; also returns the unscaled sum for logging."""

    @staticmethod
    def forward(ctx, weights, scale, *terms):
        ctx.weights, ctx.scale = tuple(float(w) for w in weights), float(scale)
        ts = [t.detach().reshape(()) for t in terms]
        total = torch.empty((), dtype=torch.float32, device=ts[0].device)
        scaled = ops.scalar_combine(ts, ctx.weights, ctx.scale, out_unscaled=total)
        ctx.mark_non_differentiable(total)
        ctx.set_materialize_grads(False)
        return scaled, total

    @staticmethod
    def backward(ctx, g, _g_total):
        if g is None:
            return (None, None) + (None,) * len(ctx.weights)
        fan = ops.scalar_fanout(g.reshape(1), ctx.weights, ctx.scale)
        return (None, None) + tuple(fan[i] for i in range(len(ctx.weights)))


def combine_losses(terms: Sequence[torch.Tensor], weights: Sequence[float], scale: float = 1.0):
    """(scale * sum_i weights[i] * terms[i]  [differentiable], the unscaled sum [detached]) for 0-dim fp32 CUDA
    loss terms - the step glue `loss = L_dino + w_g * L_gram (+ ...); loss / accum`."""
    for t in terms:
        if not (t.is_cuda and t.dtype == torch.float32 and t.numel() == 1):
            raise ValueError("combine_losses: 0-dim fp32 CUDA tensors expected")
    return _CombineLosses.apply(tuple(weights), scale, *terms)


class _KoLeo(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, eps):
        loss, saved = ops.koleo_fwd(z.detach(), eps)
        ctx.save_for_backward(z, *saved)
        ctx.eps = eps
        return loss

    @staticmethod
    def backward(ctx, g):
        z, inv, nn_idx, dist = ctx.saved_tensors
        return ops.koleo_bwd(z.detach(), (inv, nn_idx, dist), ctx.eps, g), None


class KoLeoLoss(nn.Module):
    """Kozachenko-Leonenko entropy regulariser - drop-in for scripts/phase5_big_run.py:742-773
    (`koleo_loss_fn(student_out)`, weight `--koleo-weight`, wired at :1764-1766).  The cosine Gram matrix
    runs on the tcgen05 split-K GEMM and only ranks neighbours; nearest-neighbour distances are
    recomputed exactly in fp32.  Up to 1024 rows; data parallel: local rows only (each rank regularises
    its own crops, like the single-device reference does for its batch)."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, student_output: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
        if not student_output.is_cuda:
            raise _ext.DinoxError("dinox_b200: CUDA tensors required (no CPU fallback)")
        if student_output.dim() != 2:
            raise ValueError(f"expected (rows, K) head outputs, got {tuple(student_output.shape)}")
        z = student_output if student_output.stride(1) == 1 else student_output.contiguous()
        return _KoLeo.apply(z, float(eps))

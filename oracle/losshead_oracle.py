"""CPU oracle for the DINO-X self-distillation loss head.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32) restatement of the reference algorithm for the
hot path named in BASELINE.json / SURVEY.md section 8.  It is the *checker* for the CUDA
kernels in ``dinox_b200/csrc``; nothing under ``dinox_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs use it.

Parity pin
----------
Rows a1-a10 (what the reference implements) are pinned against the reference itself:
``oracle/gen_golden.py`` imports ``/root/reference`` in the build container, runs the
reference classes on seeded inputs and stores inputs+outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors bit-for-bit-ish
(<= 1e-6 relative, same torch build).  Rows E1-E3 (multi-crop CE, Sinkhorn-Knopp, iBOT)
do NOT exist in the reference: **parity unpinned** for them.  They follow the public
DINO-v1 / DINOv2 / iBOT formulations and are tied back to the reference only by the
degenerate-case identities tested in ``tests/test_oracle_golden.py`` (n_local=0,
teacher_mode="center", no iBOT  ==  reference).

Reference citations are relative to /root/reference.

dtype policies
--------------
``policy="fp32"``  : every op in fp32 (reference without ``--amp``).
``policy="bf16"``  : operands of the dense contractions (head Linear layers, Gram bmm and
                     their backward GEMMs) are rounded to bf16, accumulation is fp32,
                     everything row-wise stays fp32.  This is the reference's CUDA-autocast
                     behaviour (scripts/phase5_big_run.py:1717) with ONE deliberate
                     difference: prototype logits and Gram entries are *not* rounded to
                     bf16 after the contraction (the CUDA path keeps them in fp32 on chip),
                     so this policy is at least as accurate as the as-run reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

__all__ = [
    "bf16_round", "head_forward", "dino_loss_reference", "center_update", "multicrop_dino_loss",
    "sinkhorn_knopp", "ibot_patch_loss", "gram_matrix", "gram_anchoring_loss", "ema_update",
    "entropy_diagnostics", "koleo_loss", "LossHeadOracle", "HeadParams",
]


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16 and back to fp32 (operand rounding of a bf16 MMA)."""
    return x.to(torch.bfloat16).to(torch.float32)


# ----------------------------------------------------------------------------------------------
# a1  projection head   zoo/arch.py:252-256 (Linear(D,D) -> GELU(erf) -> Linear(D,K)); :258-261
# ----------------------------------------------------------------------------------------------
@dataclass
class HeadParams:
    w1: torch.Tensor  # (D, D)   head.0.weight
    b1: torch.Tensor  # (D,)     head.0.bias
    w2: torch.Tensor  # (K, D)   head.2.weight
    b2: torch.Tensor  # (K,)     head.2.bias

    def tensors(self):
        return [self.w1, self.b1, self.w2, self.b2]


class _RoundBf16STE(torch.autograd.Function):
    """bf16 operand rounding in forward; in backward the incoming gradient is rounded to bf16
    as well (it is the operand of the backward GEMMs under autocast)."""

    @staticmethod
    def forward(ctx, x):
        return bf16_round(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGradBf16(torch.autograd.Function):
    """Identity in forward; rounds the gradient flowing back to bf16 (the dlogits / dGram operand
    of the backward tensor-core GEMMs)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return bf16_round(g)


def _linear(x, w, b, policy: str):
    if policy == "bf16":
        x = _RoundBf16STE.apply(x)
        w = _RoundBf16STE.apply(w)
        y = F.linear(x, w, None)
        y = _RoundGradBf16.apply(y)
        return y + b
    return F.linear(x, w, b)


def head_forward(x: torch.Tensor, p: HeadParams, policy: str = "fp32") -> torch.Tensor:
    """z = W2 . GELU(W1 . x + b1) + b2   (zoo/arch.py:252-256; exact-erf GELU = nn.GELU())."""
    a = _linear(x, p.w1, p.b1, policy)
    h = F.gelu(a)  # erf form
    return _linear(h, p.w2, p.b2, policy)


# ----------------------------------------------------------------------------------------------
# a2-a5  DINOLoss.forward / update_center      scripts/phase5_big_run.py:679-720
# ----------------------------------------------------------------------------------------------
def center_update(center: torch.Tensor, teacher_out: torch.Tensor, momentum: float,
                  global_rows: Optional[int] = None, col_sum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """c <- m*c + (1-m)*mean_rows(t)   (scripts/phase5_big_run.py:686-690).

    ``col_sum``/``global_rows`` let a data-parallel caller pass the all-reduced column sum
    (reduces to the reference at world=1)."""
    if col_sum is None:
        col_sum = teacher_out.float().sum(dim=0, keepdim=True)
        global_rows = teacher_out.shape[0]
    batch_center = col_sum.reshape(1, -1) / float(global_rows)
    return center * momentum + batch_center * (1.0 - momentum)


def dino_loss_reference(student_out, teacher_out, center, student_temp, teacher_temp):
    """Verbatim arithmetic order of DINOLoss.forward for 2 global views
    (scripts/phase5_big_run.py:703-717).  Returns the loss only (center update is separate)."""
    teacher_prob = F.softmax((teacher_out - center) / teacher_temp, dim=-1)
    student_log_prob = F.log_softmax(student_out / student_temp, dim=-1)
    B = teacher_out.shape[0] // 2
    t1, t2 = teacher_prob[:B], teacher_prob[B:]
    s1, s2 = student_log_prob[:B], student_log_prob[B:]
    loss1 = -torch.sum(t1 * s2, dim=-1).mean()
    loss2 = -torch.sum(t2 * s1, dim=-1).mean()
    return (loss1 + loss2) / 2.0


# ----------------------------------------------------------------------------------------------
# E2  Sinkhorn-Knopp teacher (EXTENSION, parity unpinned; DINOv2 public formulation;
#     motivated by docs/why-batchsize256.md:8-17 which gives no formula)
# ----------------------------------------------------------------------------------------------
def sinkhorn_knopp(teacher_out: torch.Tensor, teacher_temp: float, n_iterations: int = 3,
                   row_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Q = exp(t/tau)^T; Q/=sum(Q); repeat n_it x { Q/=rowsum (per prototype); Q/=K;
    Q/=colsum (per sample); Q/=B }; Q*=B.  ``teacher_out`` holds the GLOBAL batch (a
    data-parallel run all-reduces the per-prototype sums, so concatenating the ranks' rows is
    the definition of the distributed result).  Evaluated in the log domain (a common shift
    cancels in the first normalisation, so this is the same algebra without under/overflow).  Returns (rows, K) with rows summing to 1."""
    logq = (teacher_out.float() / teacher_temp).t()  # (K, Bg), log domain: exact same algebra,
    K, Bg = logq.shape                                # but rows far below the global max cannot
    logq = logq - torch.logsumexp(logq.reshape(-1), 0)  # underflow to 0/0
    for _ in range(n_iterations):
        logq = logq - torch.logsumexp(logq, dim=1, keepdim=True) - math.log(K)
        logq = logq - torch.logsumexp(logq, dim=0, keepdim=True) - math.log(Bg)
    logq = logq + math.log(Bg)
    return torch.exp(logq).t()


def teacher_probs(teacher_out, center, teacher_temp, teacher_mode="center", sk_iters=3):
    if teacher_mode == "center":
        return F.softmax((teacher_out.float() - center) / teacher_temp, dim=-1)
    if teacher_mode == "sinkhorn":
        return sinkhorn_knopp(teacher_out, teacher_temp, sk_iters)
    raise ValueError(f"unknown teacher_mode {teacher_mode!r}")


# ----------------------------------------------------------------------------------------------
# E1  multi-crop cross-entropy (EXTENSION, parity unpinned; DINO-v1 public formulation).
#     Row order is view-major (all B rows of view 0, then view 1, ...) as produced by
#     torch.cat(views) at scripts/phase5_big_run.py:1711.
# ----------------------------------------------------------------------------------------------
def multicrop_dino_loss(student_out, teacher_out, center, student_temp, teacher_temp,
                        n_global: int = 2, n_local: int = 0, teacher_mode: str = "center",
                        sk_iters: int = 3, teacher_prob: Optional[torch.Tensor] = None):
    """L = 1/n_terms * sum_{iq<Vg} sum_{v != iq} mean_b( -sum_k q[iq,b,k] * logp[v,b,k] ),
    n_terms = Vg*V - Vg.  With V = Vg = 2 this is scripts/phase5_big_run.py:711-717."""
    V = n_global + n_local
    Bt = teacher_out.shape[0]
    assert Bt % n_global == 0
    B = Bt // n_global
    assert student_out.shape[0] == B * V, (student_out.shape, B, V)
    q = teacher_prob if teacher_prob is not None else teacher_probs(
        teacher_out, center, teacher_temp, teacher_mode, sk_iters)
    logp = F.log_softmax(student_out.float() / student_temp, dim=-1)
    q = q.view(n_global, B, -1)
    logp = logp.view(V, B, -1)
    total = student_out.new_zeros((), dtype=torch.float32)
    n_terms = 0
    for iq in range(n_global):
        for v in range(V):
            if v == iq:
                continue
            total = total + (-(q[iq] * logp[v]).sum(dim=-1)).mean()
            n_terms += 1
    return total / n_terms


# ----------------------------------------------------------------------------------------------
# E3  iBOT masked-patch term (EXTENSION, parity unpinned; DINOv2 public formulation)
# ----------------------------------------------------------------------------------------------
def ibot_patch_loss(student_patch_out, teacher_patch_out, center_patch, student_temp, teacher_temp,
                    masks_weight: torch.Tensor, n_images: int, teacher_mode: str = "center",
                    sk_iters: int = 3, teacher_prob: Optional[torch.Tensor] = None):
    """student/teacher_patch_out: (Mm, K) logits of the SAME masked positions.
    loss = sum_m w_m * ( -sum_k q[m,k] logp[m,k] ) / n_images, w_m = 1/n_masked(image(m))."""
    q = teacher_prob if teacher_prob is not None else teacher_probs(
        teacher_patch_out, center_patch, teacher_temp, teacher_mode, sk_iters)
    logp = F.log_softmax(student_patch_out.float() / student_temp, dim=-1)
    per_tok = -(q * logp).sum(dim=-1)
    return (per_tok * masks_weight).sum() / float(n_images)


# ----------------------------------------------------------------------------------------------
# a6-a7  Gram anchoring     scripts/phase5_big_run.py:723-739
# ----------------------------------------------------------------------------------------------
def gram_matrix(feats: torch.Tensor, policy: str = "fp32") -> torch.Tensor:
    """Xn = X / max(||X||, 1e-12);  G = Xn Xn^T per image  (scripts/phase5_big_run.py:723-728)."""
    xn = F.normalize(feats.float(), p=2, dim=-1)
    if policy == "bf16":
        xn = _RoundBf16STE.apply(xn)
        g = torch.bmm(xn, xn.transpose(1, 2))
        return _RoundGradBf16.apply(g)
    return torch.bmm(xn, xn.transpose(1, 2))


def gram_anchoring_loss(student_feats, teacher_feats, policy: str = "fp32"):
    """mse_loss(G(student[:,1:]), G(teacher[:,1:]))  (scripts/phase5_big_run.py:731-739).
    feats[:,1:] = patches AND the 4 register tokens (zoo/arch.py:219-229)."""
    gs = gram_matrix(student_feats[:, 1:], policy)
    gt = gram_matrix(teacher_feats[:, 1:], policy)
    return F.mse_loss(gs, gt)


# ----------------------------------------------------------------------------------------------
# a8  EMA teacher update   scripts/phase5_big_run.py:1798-1802 ; scripts/phase3_micro_run.py:152-155
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def ema_update(teacher_params: Sequence[torch.Tensor], student_params: Sequence[torch.Tensor], m: float) -> None:
    """p_t <- m*p_t + (1-m)*p_s, in place, reference op order (mul_ then add_ with alpha)."""
    for p_t, p_s in zip(teacher_params, student_params):
        p_t.mul_(m).add_(p_s, alpha=1.0 - m)


# ----------------------------------------------------------------------------------------------
# a10  entropy diagnostics   scripts/phase5_big_run.py:1843-1853
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def entropy_diagnostics(student_out, teacher_out, center, student_temp, teacher_temp):
    t_logits = (teacher_out.float() - center.float()) / teacher_temp
    s_logits = student_out.float() / student_temp
    t_ent = -(F.softmax(t_logits, -1) * F.log_softmax(t_logits, -1)).sum(-1).mean()
    s_ent = -(F.softmax(s_logits, -1) * F.log_softmax(s_logits, -1)).sum(-1).mean()
    return t_ent, s_ent


# ----------------------------------------------------------------------------------------------
# a11  KoLeo (SURVEY 8f "next #1")   scripts/phase5_big_run.py:742-773
# ----------------------------------------------------------------------------------------------
def koleo_loss(student_output: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    x = F.normalize(student_output.float(), p=2, dim=-1)
    pdist = torch.cdist(x, x, p=2)
    pdist = pdist + torch.eye(x.shape[0]) * 1e9
    min_dist, _ = pdist.min(dim=1)
    return -torch.log(min_dist + eps).mean()


# ----------------------------------------------------------------------------------------------
# Whole micro-step (a9 step glue, scripts/phase5_big_run.py:1738-1772) on pre-extracted features
# ----------------------------------------------------------------------------------------------
class LossHeadOracle:
    """One loss-head micro-step on CPU.

    Inputs are backbone outputs (the backbone is out of scope):
      student_cls (B*V, D)  view-major CLS rows of all V = Vg+Vl crops
      teacher_cls (B*Vg, D) CLS rows of the global crops
      student_tok / teacher_tok (B*Vg, T, D) full token tensors of the global crops (Gram)
      student_patch / teacher_patch (Mm, D) masked-position patch rows (iBOT), masks_weight (Mm,)
    State: center (1,K) fp32 (DINOLoss buffer, :684), center_patch (1,K) fp32 (extension).
    """

    def __init__(self, student: HeadParams, teacher: HeadParams, out_dim: int,
                 center_momentum: float = 0.999, n_global: int = 2, n_local: int = 0,
                 teacher_mode: str = "center", sk_iters: int = 3, gram_weight: float = 1.0,
                 ibot_weight: float = 1.0, policy: str = "fp32", patch_teacher_mode: Optional[str] = None):
        # patch_teacher_mode: normalisation of the iBOT (masked-patch) teacher rows; None = same as teacher_mode.
        # "center" with teacher_mode="sinkhorn" is what the CUDA path implements: Sinkhorn-Knopp on the CLS rows,
        # softmax-centring with the patch centre (kept updated) on the patch rows.
        self.patch_teacher_mode = patch_teacher_mode or teacher_mode
        self.student, self.teacher = student, teacher
        self.center = torch.zeros(1, out_dim)
        self.center_patch = torch.zeros(1, out_dim)
        self.center_momentum = center_momentum
        self.n_global, self.n_local = n_global, n_local
        self.teacher_mode, self.sk_iters = teacher_mode, sk_iters
        self.gram_weight, self.ibot_weight = gram_weight, ibot_weight
        self.policy = policy

    def step(self, student_cls, teacher_cls, student_temp, teacher_temp,
             student_tok=None, teacher_tok=None, student_patch=None, teacher_patch=None,
             masks_weight=None, accum: int = 1, update_center: bool = True):
        out = {}
        s_out = head_forward(student_cls, self.student, self.policy)
        with torch.no_grad():
            t_out = head_forward(teacher_cls, self.teacher, self.policy)
        loss_dino = multicrop_dino_loss(s_out, t_out, self.center, student_temp, teacher_temp,
                                        self.n_global, self.n_local, self.teacher_mode, self.sk_iters)
        loss = loss_dino
        out["loss_dino"] = loss_dino.detach()
        if student_patch is not None:
            sp_out = head_forward(student_patch, self.student, self.policy)
            with torch.no_grad():
                tp_out = head_forward(teacher_patch, self.teacher, self.policy)
            n_images = teacher_cls.shape[0]
            loss_ibot = ibot_patch_loss(sp_out, tp_out, self.center_patch, student_temp, teacher_temp,
                                        masks_weight, n_images, self.patch_teacher_mode, self.sk_iters)
            loss = loss + self.ibot_weight * loss_ibot
            out["loss_ibot"] = loss_ibot.detach()
        if student_tok is not None:
            loss_gram = gram_anchoring_loss(student_tok, teacher_tok.detach(), self.policy)
            loss = loss + self.gram_weight * loss_gram
            out["loss_gram"] = loss_gram.detach()
        out["loss"] = loss.detach()
        (loss / accum).backward()
        if update_center:
            with torch.no_grad():
                if self.teacher_mode == "center":
                    self.center = center_update(self.center, t_out, self.center_momentum)
                if student_patch is not None and self.patch_teacher_mode == "center":
                    self.center_patch = center_update(self.center_patch, tp_out, self.center_momentum)
        return out

    def ema(self, student_all: Sequence[torch.Tensor], teacher_all: Sequence[torch.Tensor], m: float):
        ema_update(teacher_all, student_all, m)

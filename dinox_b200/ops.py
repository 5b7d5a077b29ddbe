"""Thin torch-tensor wrappers over the C ABI (include/dinox_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every computation below is one
call into libdinox_b200.so with raw pointers.  Nothing in this file computes on the CPU or
through ATen - if the library is missing or the device is not a B200 the call raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _ext

DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
LOG2E = 1.4426950408889634


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _ext.DinoxError("dinox_b200 ops need CUDA tensors (there is no CPU fallback)")


def _rowmajor(t: torch.Tensor) -> int:
    """leading dimension of a 2-D tensor whose last dim is contiguous"""
    if t.dim() != 2 or t.stride(1) != 1:
        raise _ext.DinoxError(f"expected a 2-D tensor with contiguous rows, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def launch_count() -> int:
    return int(_ext.lib().dinox_launch_count())


def launch_count_reset() -> None:
    _ext.lib().dinox_launch_count_reset()


# ------------------------------------------------------------------------------------------------
# a8 EMA
# ------------------------------------------------------------------------------------------------
class EmaPlan:
    """Multi-tensor EMA plan over (student, teacher) parameter pairs (fp32, same shapes)."""

    def __init__(self, student: Sequence[torch.Tensor], teacher: Sequence[torch.Tensor]):
        student, teacher = list(student), list(teacher)
        if len(student) != len(teacher):
            raise ValueError("student/teacher parameter lists differ in length")
        _chk_cuda(*student, *teacher)
        for s, t in zip(student, teacher):
            if s.shape != t.shape or s.dtype != torch.float32 or t.dtype != torch.float32:
                raise ValueError("EMA needs fp32 parameters of identical shapes")
            if not (s.is_contiguous() and t.is_contiguous()):
                raise ValueError("EMA needs contiguous parameters")
        n = len(student)
        ps = (ctypes.c_void_p * n)(*[s.data_ptr() for s in student])
        pt = (ctypes.c_void_p * n)(*[t.data_ptr() for t in teacher])
        ne = (ctypes.c_int64 * n)(*[s.numel() for s in student])
        self._h = ctypes.c_void_p()
        self._keys = [(s.data_ptr(), t.data_ptr(), s.numel()) for s, t in zip(student, teacher)]
        _ext.call("dinox_ema_plan_create", ps, pt, ne, n, ctypes.byref(self._h))
        self.numel = int(_ext.lib().dinox_ema_plan_numel(self._h))

    def matches(self, student, teacher) -> bool:
        keys = [(s.data_ptr(), t.data_ptr(), s.numel()) for s, t in zip(student, teacher)]
        return keys == self._keys

    def apply(self, m: float) -> None:
        # alpha = 1.0 - m is evaluated in double like the reference's python expression
        _ext.call("dinox_ema_apply", self._h, float(m), float(1.0 - float(m)), _stream())

    def __del__(self):
        try:
            if self._h:
                _ext.lib().dinox_ema_plan_destroy(self._h)
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# row / column statistics on materialised logits
# ------------------------------------------------------------------------------------------------
def rows_lse(x: torch.Tensor, inv_tau: float, colbias: Optional[torch.Tensor] = None,
             want_entropy: bool = False):
    _chk_cuda(x, colbias)
    ld = _rowmajor(x)
    rows, K = x.shape
    lse = torch.empty(rows, dtype=torch.float32, device=x.device)
    ent = torch.empty(rows, dtype=torch.float32, device=x.device) if want_entropy else None
    _ext.call("dinox_rows_lse", _p(x), DT[x.dtype], rows, K, ld, float(inv_tau), _p(colbias), _p(lse), _p(ent), _stream())
    return (lse, ent) if want_entropy else lse


def cols_lse(x: torch.Tensor, inv_tau: float, rowbias: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk_cuda(x, rowbias)
    ld = _rowmajor(x)
    rows, K = x.shape
    out = torch.empty(K, dtype=torch.float32, device=x.device)
    _ext.call("dinox_cols_lse", _p(x), DT[x.dtype], rows, K, ld, float(inv_tau), _p(rowbias), _p(out), _stream())
    return out


def lse_combine(gathered: torch.Tensor, add: float = 0.0) -> torch.Tensor:
    world, K = gathered.shape
    out = torch.empty(K, dtype=torch.float32, device=gathered.device)
    _ext.call("dinox_lse_combine", _p(gathered), world, K, float(add), _p(out), _stream())
    return out


def cols_sum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk_cuda(x)
    ld = _rowmajor(x)
    rows, K = x.shape
    if out is None:
        out = torch.empty(K, dtype=torch.float32, device=x.device)
    assert out.dtype == torch.float32 and out.numel() == K and out.is_contiguous()
    if rows >= 1024 and K <= 8192:
        # tall and skinny: K/128 CTAs would each walk every row - sum 64-row chunks first (fixed order)
        chunk = 64
        part = torch.empty((rows + chunk - 1) // chunk, K, dtype=torch.float32, device=x.device)
        _ext.call("dinox_cols_sum_chunked", _p(x), DT[x.dtype], rows, K, ld, chunk, _p(part), _stream())
        _ext.call("dinox_cols_sum", _p(part), DT[torch.float32], part.shape[0], K, K, _p(out), _stream())
        return out
    _ext.call("dinox_cols_sum", _p(x), DT[x.dtype], rows, K, ld, _p(out), _stream())
    return out


def cols_sum_axpy_(x: torch.Tensor, out: torch.Tensor, accumulate: bool, scale: float = 1.0,
                   scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (+)= scale * scale_dev * x.sum(0) in one launch (fixed summation order)."""
    _chk_cuda(x, out)
    rows, K = x.shape
    assert out.dtype == torch.float32 and out.numel() == K and out.is_contiguous()
    _ext.call("dinox_cols_sum_axpy", _p(x), DT[x.dtype], rows, K, _rowmajor(x), float(scale), _p(scale_dev), _p(out),
              int(accumulate), _stream())
    return out


def center_ema_(center: torch.Tensor, colsum: torch.Tensor, global_rows: int, momentum: float) -> None:
    _chk_cuda(center, colsum)
    K = center.numel()
    _ext.call("dinox_center_ema", _p(center), _p(colsum), float(1.0 / global_rows), float(momentum), K, _stream())


def axpb(a: torch.Tensor, alpha: float, beta: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty_like(a) if out is None else out
    _ext.call("dinox_axpb", _p(a), float(alpha), float(beta), _p(out), a.numel(), _stream())
    return out


def ce_fwd(student, teacher, groups: int, V: int, Vg: int, inv_tau_s: float, inv_tau_t: float,
           colbias_t, rowbias_t, lse_s, group_w, norm: float, exclude_same: bool) -> torch.Tensor:
    _chk_cuda(student, teacher)
    K = student.shape[1]
    ws = torch.empty(int(_ext.lib().dinox_ce_workspace_bytes(groups, K)), dtype=torch.uint8, device=student.device)
    loss = torch.empty((), dtype=torch.float32, device=student.device)
    _ext.call("dinox_ce_fwd", _p(student), DT[student.dtype], _p(teacher), DT[teacher.dtype], groups, V, Vg, K,
              _rowmajor(student), _rowmajor(teacher), float(inv_tau_s), float(inv_tau_t), _p(colbias_t),
              _p(rowbias_t), _p(lse_s), _p(group_w), float(norm), int(exclude_same), _p(loss), _p(ws), _stream())
    return loss


def ce_onepass_max_views() -> int:
    return int(_ext.lib().dinox_ce_onepass_max_views())


def ce_fwd_onepass(student, teacher, groups: int, V: int, Vg: int, inv_tau_s: float, inv_tau_t: float,
                   colbias_t, group_w, norm: float, exclude_same: bool):
    """Cross-entropy forward with a softmax-centred teacher in ONE pass over the logits.  Returns
    (loss, lse_s (V*groups), rowbias_t (Vg*groups)); the two LSE vectors are what `ce_bwd` needs."""
    _chk_cuda(student, teacher, colbias_t, group_w)
    K = student.shape[1]
    dev = student.device
    ws = torch.empty(int(_ext.lib().dinox_ce_onepass_workspace_bytes(groups, V, Vg, K)), dtype=torch.uint8, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    lse_s = torch.empty(V * groups, dtype=torch.float32, device=dev)
    rowbias_t = torch.empty(Vg * groups, dtype=torch.float32, device=dev)
    _ext.call("dinox_ce_fwd_onepass", _p(student), DT[student.dtype], _p(teacher), DT[teacher.dtype], groups, V, Vg, K,
              _rowmajor(student), _rowmajor(teacher), float(inv_tau_s), float(inv_tau_t), _p(colbias_t), _p(group_w),
              float(norm), int(exclude_same), _p(loss), _p(lse_s), _p(rowbias_t), _p(ws), _stream())
    return loss, lse_s, rowbias_t


def ce_bwd(student, teacher, groups: int, V: int, Vg: int, inv_tau_s: float, inv_tau_t: float,
           colbias_t, rowbias_t, lse_s, group_w, norm: float, exclude_same: bool,
           upstream: torch.Tensor) -> torch.Tensor:
    K = student.shape[1]
    grad = torch.empty_like(student, memory_format=torch.contiguous_format)
    up = upstream.to(torch.float32).reshape(1).contiguous()
    _ext.call("dinox_ce_bwd", _p(student), DT[student.dtype], _p(teacher), DT[teacher.dtype], groups, V, Vg, K,
              _rowmajor(student), _rowmajor(teacher), float(inv_tau_s), float(inv_tau_t), _p(colbias_t),
              _p(rowbias_t), _p(lse_s), _p(group_w), float(norm), int(exclude_same), _p(up), _p(grad),
              _rowmajor(grad), _stream())
    return grad


# ------------------------------------------------------------------------------------------------
# tcgen05 GEMMs
# ------------------------------------------------------------------------------------------------
def gemm_bf16(a: torch.Tensor, b: torch.Tensor, *, a_mn_major: bool = False, b_mn_major: bool = False,
              out: Optional[torch.Tensor] = None, out_dtype=torch.float32, accumulate: bool = False,
              alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None,
              bias_n: Optional[torch.Tensor] = None, m_fastest: bool = True) -> torch.Tensor:
    """C[M,N] (+)= alpha * A @ B^T + bias_n.   a: (M,K) [or (K,M) if a_mn_major], b: (N,K) [or (K,N)]."""
    _chk_cuda(a, b, out, bias_n, alpha_dev)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise _ext.DinoxError("gemm_bf16 needs bf16 operands")
    M, Ka = (a.shape[1], a.shape[0]) if a_mn_major else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else b.shape
    if Ka != Kb:
        raise _ext.DinoxError(f"gemm_bf16: reduction dims differ ({Ka} vs {Kb})")
    if out is None:
        if accumulate:
            raise _ext.DinoxError("accumulate=True needs an output tensor")
        q = 16 // torch.empty((), dtype=out_dtype).element_size()      # 16-byte row pitch for the TMA store
        out = torch.empty(M, (N + q - 1) // q * q, dtype=out_dtype, device=a.device)[:, :N]
    _ext.call("dinox_gemm_bf16", _p(a), _p(b), _p(out), M, N, Ka, _rowmajor(a), _rowmajor(b), _rowmajor(out),
              int(a_mn_major), int(b_mn_major), DT[out.dtype], int(accumulate), float(alpha), _p(alpha_dev),
              _p(bias_n), int(m_fastest), _stream())
    return out


def gemm_bf16_balanced(a: torch.Tensor, b: torch.Tensor, *, tag: str, a_mn_major: bool = False, b_mn_major: bool = False,
                       out: Optional[torch.Tensor] = None, accumulate: bool = False, alpha: float = 1.0,
                       alpha_dev: Optional[torch.Tensor] = None, bias_n: Optional[torch.Tensor] = None,
                       m_fastest: bool = True) -> torch.Tensor:
    """gemm_bf16 with fp32 output and the balanced schedule (tiles of a partial wave cut along K, parts accumulated
    in a fixed order).  `tag` names the call site: its ordering counters live in a persistent zeroed workspace."""
    _chk_cuda(a, b, out, bias_n, alpha_dev)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise _ext.DinoxError("gemm_bf16_balanced needs bf16 operands")
    M, Ka = (a.shape[1], a.shape[0]) if a_mn_major else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else b.shape
    if Ka != Kb:
        raise _ext.DinoxError(f"gemm_bf16_balanced: reduction dims differ ({Ka} vs {Kb})")
    if out is None:
        if accumulate:
            raise _ext.DinoxError("accumulate=True needs an output tensor")
        out = torch.empty(M, (N + 3) // 4 * 4, dtype=torch.float32, device=a.device)[:, :N]
    assert out.dtype == torch.float32
    flags = zeroed_workspace("gemm_balanced:" + tag, int(_ext.lib().dinox_gemm_bf16_balanced_workspace_bytes(M, N)), a.device)
    _ext.call("dinox_gemm_bf16_balanced", _p(a), _p(b), _p(out), M, N, Ka, _rowmajor(a), _rowmajor(b), _rowmajor(out),
              int(a_mn_major), int(b_mn_major), int(accumulate), float(alpha), _p(alpha_dev), _p(bias_n), int(m_fastest),
              _p(flags), _stream())
    return out


def gemm_bf16_reduce_scatter(a: torch.Tensor, b: torch.Tensor, shard_ptrs: Sequence[int], rows_per_owner: int, ldc: int,
                             *, a_mn_major: bool = False, b_mn_major: bool = False, alpha: float = 1.0,
                             alpha_dev: Optional[torch.Tensor] = None, add_local: Optional[torch.Tensor] = None,
                             add_scale: float = 1.0) -> None:
    """alpha * A @ B^T reduce-added, tile by tile, into the row shards named by `shard_ptrs` (device addresses, one
    per data-parallel rank, possibly peer-mapped): the reduce-scatter happens in the GEMM epilogue.  `add_local`
    ((M, N) fp32, times add_scale) rides along: the gradient accumulated locally over the earlier micro-steps."""
    _chk_cuda(a, b, alpha_dev)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise _ext.DinoxError("gemm_bf16_reduce_scatter needs bf16 operands")
    M, Ka = (a.shape[1], a.shape[0]) if a_mn_major else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else b.shape
    if Ka != Kb:
        raise _ext.DinoxError(f"gemm_bf16_reduce_scatter: reduction dims differ ({Ka} vs {Kb})")
    n = len(shard_ptrs)
    ptrs = (ctypes.c_void_p * n)(*[int(p) for p in shard_ptrs])
    if add_local is not None:
        assert add_local.dtype == torch.float32 and tuple(add_local.shape) == (M, N) and add_local.stride(1) == 1
    _ext.call("dinox_gemm_bf16_reduce_scatter", _p(a), _p(b), ptrs, n, int(rows_per_owner), M, N, Ka, _rowmajor(a),
              _rowmajor(b), int(ldc), int(a_mn_major), int(b_mn_major), float(alpha), _p(alpha_dev), _p(add_local),
              0 if add_local is None else _rowmajor(add_local), float(add_scale), _stream())


def gemm_splitk_plan(M: int, N: int, K: int) -> int:
    return int(_ext.lib().dinox_gemm_splitk_plan(M, N, K))


def gemm_bf16_splitk(a: torch.Tensor, b: torch.Tensor, *, a_mn_major: bool = False, b_mn_major: bool = False,
                     splits: Optional[int] = None, alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None,
                     m_fastest: bool = True) -> torch.Tensor:
    """Split-K GEMM: returns the (splits, M, N) fp32 partial slabs (sum over dim 0 = alpha * A @ B^T).
    splits=None asks the library for the count that fills whole waves of the persistent grid."""
    _chk_cuda(a, b, alpha_dev)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise _ext.DinoxError("gemm_bf16_splitk needs bf16 operands")
    M, Ka = (a.shape[1], a.shape[0]) if a_mn_major else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else b.shape
    if Ka != Kb:
        raise _ext.DinoxError(f"gemm_bf16_splitk: reduction dims differ ({Ka} vs {Kb})")
    if splits is None:
        splits = gemm_splitk_plan(M, N, Ka)
    out = torch.empty(splits, M, N, dtype=torch.float32, device=a.device)
    _ext.call("dinox_gemm_bf16_splitk", _p(a), _p(b), _p(out), M, N, Ka, _rowmajor(a), _rowmajor(b), N, M * N, splits,
              int(a_mn_major), int(b_mn_major), float(alpha), _p(alpha_dev), int(m_fastest), _stream())
    return out


def head_stats(h: torch.Tensor, w2: torch.Tensor, inv_tau: float, col2: Optional[torch.Tensor] = None,
               want_nat: bool = True, want_log2: bool = True, out_log2: Optional[torch.Tensor] = None):
    """Row-wise LSE of (h @ w2^T)*inv_tau + col2/log2e without materialising the logits."""
    _chk_cuda(h, w2, col2)
    rows, D = h.shape
    K = w2.shape[0]
    ws = torch.empty(int(_ext.lib().dinox_head_stats_workspace_bytes(rows, K)), dtype=torch.uint8, device=h.device)
    nat = torch.empty(rows, dtype=torch.float32, device=h.device) if want_nat else None
    l2 = out_log2 if out_log2 is not None else (
        torch.empty(rows, dtype=torch.float32, device=h.device) if want_log2 else None)
    _ext.call("dinox_head_stats", _p(h), _p(w2), rows, K, D, _rowmajor(h), _rowmajor(w2), float(inv_tau), _p(col2),
              _p(nat), _p(l2), _p(ws), _stream())
    return nat, l2


def head_grad(w2s, w2t, hs_e, ht_e, inv_tau_s, inv_tau_t, cs2, ct2, ct2_alt, alt_from, lse2_e, rb2_e, cw_e,
              loss_out: torch.Tensor, loss_accumulate: bool = False, want_db2: bool = True,
              gt: Optional[torch.Tensor] = None):
    """Pass 2.  loss_out: 2 fp32 ([0] entries < alt_from, [1] the rest).  Returns (Gt (K, E_pad) bf16,
    db2_partial or None)."""
    assert loss_out.numel() >= 2
    _chk_cuda(w2s, w2t, hs_e, ht_e)
    K, D = w2s.shape
    E = hs_e.shape[0]
    e_pad = (E + 127) // 128 * 128
    if gt is None:
        gt = torch.empty(K, e_pad, dtype=torch.bfloat16, device=w2s.device)
    db2p = (torch.empty(int(_ext.lib().dinox_head_grad_db2_rows(E)), K, dtype=torch.float32, device=w2s.device)
            if want_db2 else None)
    ws = torch.empty(int(_ext.lib().dinox_head_grad_workspace_bytes(K, E)), dtype=torch.uint8, device=w2s.device)
    _ext.call("dinox_head_grad", _p(w2s), _p(w2t), _p(hs_e), _p(ht_e), K, D, E, _rowmajor(w2s), _rowmajor(w2t),
              _rowmajor(hs_e), _rowmajor(ht_e), float(inv_tau_s), float(inv_tau_t), _p(cs2), _p(ct2), _p(ct2_alt),
              int(alt_from), _p(lse2_e), _p(rb2_e), _p(cw_e), _p(gt), _rowmajor(gt), _p(db2p), _p(loss_out),
              int(loss_accumulate), _p(ws), _stream())
    return gt, db2p


def teacher_granules_per_tile() -> int:
    """granules (prototype groups sharing one reference exponent) per prototype tile: 2, 3 or 4"""
    return int(_ext.lib().dinox_head_teacher_granules_per_tile())


def teacher_tile_cols() -> int:
    """prototypes per tile of the read-back pair (256, or 192 in the 12-epilogue-warp build)"""
    return int(_ext.lib().dinox_head_teacher_tile_cols())


def teacher_buffers(rows: int, K: int, device):
    """(qt, refs) of dinox_head_teacher: fp16 probabilities with rows padded to whole prototype tiles,
    and the per-(granule, row) maxima."""
    tile = teacher_tile_cols()
    n_tiles = (K + tile - 1) // tile
    qt = torch.empty(rows, n_tiles * tile, dtype=torch.float16, device=device)
    refs = torch.empty(teacher_granules_per_tile() * n_tiles, rows, dtype=torch.float32, device=device)
    return qt, refs


def head_teacher(h: torch.Tensor, w2: torch.Tensor, inv_tau: float, col2: Optional[torch.Tensor],
                 col2_alt: Optional[torch.Tensor] = None, alt_from_row: int = 0,
                 qt: Optional[torch.Tensor] = None, refs: Optional[torch.Tensor] = None,
                 out_log2: Optional[torch.Tensor] = None):
    """Teacher in one pass: returns (qt (rows, K_pad) fp16, refs (granules, rows) fp32, lse2 (rows)) with
    softmax(h @ w2^T * inv_tau + col2/log2e)[i, k] = qt[i, k] * 2^(refs[k // gw, i] - lse2[i]),
    gw = teacher_tile_cols() // teacher_granules_per_tile() prototypes per granule."""
    _chk_cuda(h, w2, col2, col2_alt)
    rows, D = h.shape
    K = w2.shape[0]
    if qt is None or refs is None:
        qt, refs = teacher_buffers(rows, K, h.device)
    l2 = out_log2 if out_log2 is not None else torch.empty(rows, dtype=torch.float32, device=h.device)
    ws = torch.empty(int(_ext.lib().dinox_head_teacher_workspace_bytes(rows, K)), dtype=torch.uint8, device=h.device)
    _ext.call("dinox_head_teacher", _p(h), _p(w2), rows, K, D, _rowmajor(h), _rowmajor(w2), float(inv_tau), _p(col2),
              _p(col2_alt), int(alt_from_row), _p(qt), qt.stride(0), _p(refs), refs.stride(0), None, _p(l2), _p(ws),
              _stream())
    return qt, refs, l2


_ZEROED: dict = {}


def zeroed_workspace(tag: str, nbytes: int, device) -> torch.Tensor:
    """Persistent uint8 workspace for the kernels that keep a ticket counter in it (zero before the first launch, left
    zero by every launch).  One per (tag, size, device): launches of one tag must be stream-ordered (they are - each
    tag belongs to one call site of the step); concurrent users pass their own tag."""
    dev = torch.device(device)
    key = (tag, int(nbytes), dev.index if dev.index is not None else torch.cuda.current_device())
    t = _ZEROED.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            # allocated from the capturing graph's pool the block would be recycled between replays: make it once, eagerly
            raise _ext.DinoxError(f"workspace {tag!r} must exist before graph capture (run one eager warm-up step first)")
        t = _ZEROED[key] = torch.zeros(int(nbytes), dtype=torch.uint8, device=dev)
    return t


def head_grad2(hs_e: torch.Tensor, w2s: torch.Tensor, inv_tau_s: float, cs2: torch.Tensor, lse2_e: torch.Tensor,
               cw_e: torch.Tensor, rb2_e: torch.Tensor, trow_e: torch.Tensor, qt: torch.Tensor, refs: torch.Tensor,
               alt_from: int, loss_out: torch.Tensor, loss_accumulate: bool = False, want_db2: bool = True,
               g: Optional[torch.Tensor] = None, srow_e: Optional[torch.Tensor] = None, fused_loss_sum: bool = True):
    """Pass 2 (student logits recomputed, teacher probabilities read back).  loss_out: 3 fp32 ([0] entries <
    alt_from, [1] the rest, [2] their sum).  Returns (G (E, K) bf16 = dL/dlogits per entry, db2_partial or None).
    srow_e (int32 student row per entry, -1 = padding): lse2_e is then the per-ROW student LSE and rb2_e the per-ROW
    teacher offset (indexed through srow_e / trow_e inside the kernel).  fused_loss_sum: the kernel adds up its own
    loss partials (ticket counter) instead of a follow-up launch."""
    assert loss_out.numel() >= 3 and trow_e.dtype == torch.int32 and qt.dtype == torch.float16
    assert srow_e is None or (srow_e.dtype == torch.int32 and srow_e.numel() == trow_e.numel())
    _chk_cuda(hs_e, w2s, qt, refs)
    E, D = hs_e.shape
    K = w2s.shape[0]
    if g is None:   # rows padded to 16 bytes (TMA store pitch); the view hides the padding
        g = torch.empty(E, (K + 7) // 8 * 8, dtype=torch.bfloat16, device=w2s.device)[:, :K]
    db2p = (torch.empty(int(_ext.lib().dinox_head_grad2_db2_rows(E)), K, dtype=torch.float32, device=w2s.device)
            if want_db2 else None)
    ws = torch.empty(int(_ext.lib().dinox_head_grad2_workspace_bytes(E, K)), dtype=torch.uint8, device=w2s.device)
    ticket = zeroed_workspace("head_grad2", 256, w2s.device) if fused_loss_sum else None
    _ext.call("dinox_head_grad2", _p(hs_e), _p(w2s), E, K, D, _rowmajor(hs_e), _rowmajor(w2s), float(inv_tau_s), _p(cs2),
              _p(lse2_e), _p(cw_e), _p(rb2_e), _p(trow_e), _p(srow_e), _p(qt), qt.stride(0), _p(refs), refs.stride(0),
              int(alt_from), _p(g), _rowmajor(g), _p(db2p), _p(loss_out), int(loss_accumulate), _p(ws), _p(ticket), _stream())
    return g, db2p


def axpby(x: torch.Tensor, alpha: float, y: Optional[torch.Tensor], beta: float,
          out: Optional[torch.Tensor] = None, alpha_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty_like(x) if out is None else out
    _ext.call("dinox_axpby", _p(x), float(alpha), _p(alpha_dev), _p(y), float(beta), _p(out), x.numel(), _stream())
    return out


def gemm_bf16_batched(a: torch.Tensor, b: torch.Tensor, *, a_mn_major=False, b_mn_major=False,
                      out: Optional[torch.Tensor] = None, out_dtype=torch.float32, accumulate=False, alpha=1.0,
                      alpha_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a: (B, M, K) [or (B, K, M)], b: (B, N, K) [or (B, K, N)], contiguous inner matrices."""
    _chk_cuda(a, b, out)
    Bt = a.shape[0]
    M, Ka = (a.shape[2], a.shape[1]) if a_mn_major else (a.shape[1], a.shape[2])
    N, Kb = (b.shape[2], b.shape[1]) if b_mn_major else (b.shape[1], b.shape[2])
    assert Ka == Kb and b.shape[0] == Bt and a.stride(2) == 1 and b.stride(2) == 1
    if out is None:
        out = torch.empty(Bt, M, N, dtype=out_dtype, device=a.device)
    _ext.call("dinox_gemm_bf16_batched", _p(a), _p(b), _p(out), Bt, M, N, Ka, a.stride(1), b.stride(1), out.stride(1),
              a.stride(0), b.stride(0), out.stride(0), int(a_mn_major), int(b_mn_major), DT[out.dtype], int(accumulate),
              float(alpha), _p(alpha_dev), _stream())
    return out


def normalize_tokens(feats: torch.Tensor, skip: int = 1):
    """feats (B, T, D) fp32|bf16 with contiguous last dim -> xn (B, T-skip, D) bf16, inv_norm (B, T-skip)."""
    _chk_cuda(feats)
    Bt, T, D = feats.shape
    if feats.stride(2) != 1:
        feats = feats.contiguous()
    xn = torch.empty(Bt, T - skip, D, dtype=torch.bfloat16, device=feats.device)
    inv = torch.empty(Bt, T - skip, dtype=torch.float32, device=feats.device)
    _ext.call("dinox_normalize_tokens", _p(feats), DT[feats.dtype], Bt, T, D, feats.stride(0), feats.stride(1), skip,
              _p(xn), _p(inv), _stream())
    return xn, inv


def normalize_tokens_bwd(feats: torch.Tensor, dxn: torch.Tensor, inv_norm: torch.Tensor, grad: torch.Tensor,
                         skip: int = 1, scale: float = 1.0, scale_dev: Optional[torch.Tensor] = None) -> None:
    Bt, T, D = feats.shape
    _ext.call("dinox_normalize_tokens_bwd", _p(feats), DT[feats.dtype], Bt, T, D, feats.stride(0), feats.stride(1), skip,
              _p(dxn), _p(inv_norm), _p(scale_dev), float(scale), _p(grad), grad.stride(0), grad.stride(1), _stream())


def gram_diff(xn_s: torch.Tensor, xn_t: torch.Tensor, loss_scale: float, want_delta: bool = True):
    """loss = loss_scale * sum_b ||Xs Xs^T - Xt Xt^T||_F^2 ; delta (B, T', ldd) bf16 (optional)."""
    Bt, Tp, D = xn_s.shape
    ldd = (Tp + 7) // 8 * 8
    delta = torch.empty(Bt, Tp, ldd, dtype=torch.bfloat16, device=xn_s.device) if want_delta else None
    ws = torch.empty(int(_ext.lib().dinox_gram_diff_workspace_bytes(Bt, Tp)), dtype=torch.uint8, device=xn_s.device)
    loss = torch.empty((), dtype=torch.float32, device=xn_s.device)
    _ext.call("dinox_gram_diff", _p(xn_s), _p(xn_t), Bt, Tp, D, _p(delta), ldd, float(loss_scale), _p(loss), _p(ws), _stream())
    return loss, delta


def gather_cast_bf16(src: torch.Tensor, idx: Optional[torch.Tensor], out: torch.Tensor) -> torch.Tensor:
    """out[r] = bf16(src[idx[r]]) (idx None: identity; idx < 0: zero row). src rows may be strided."""
    _chk_cuda(src, out)
    assert src.dim() == 2 and src.stride(1) == 1 and out.stride(1) == 1
    rows = out.shape[0]
    _ext.call("dinox_gather_cast_bf16", _p(src), DT[src.dtype], src.stride(0), _p(idx), rows, src.shape[1], _p(out),
              out.stride(0), _stream())
    return out


def gather_cast_bf16_2(src0: torch.Tensor, idx0: Optional[torch.Tensor], rows0: int, src1: torch.Tensor,
                       idx1: Optional[torch.Tensor], out: torch.Tensor) -> torch.Tensor:
    """out[:rows0] = bf16(src0[idx0]), out[rows0:] = bf16(src1[idx1]) in one launch when both sources are 16-byte
    vector rows of one dtype; two launches otherwise.  idx None: identity; idx < 0: zero row."""
    _chk_cuda(src0, src1, out)
    rows1 = out.shape[0] - rows0
    D = out.shape[1]
    es = src0.element_size()
    fits = (src0.dtype == src1.dtype and D % 8 == 0 and src0.stride(1) == 1 and src1.stride(1) == 1 and out.stride(1) == 1
            and all(t.data_ptr() % 16 == 0 for t in (src0, src1, out))
            and (src0.stride(0) * es) % 16 == 0 and (src1.stride(0) * es) % 16 == 0 and (out.stride(0) * 2) % 16 == 0)
    if not fits or rows0 == 0 or rows1 == 0:
        if rows0:
            gather_cast_bf16(src0, idx0, out[:rows0])
        if rows1:
            gather_cast_bf16(src1, idx1, out[rows0:])
        return out
    _ext.call("dinox_gather_cast_bf16_2", _p(src0), src0.stride(0), _p(idx0), rows0, _p(src1), src1.stride(0), _p(idx1),
              rows1, DT[src0.dtype], D, _p(out), out.stride(0), _stream())
    return out


def segment_cols_sum(x: torch.Tensor, segments: Sequence[Tuple[int, int]], out: torch.Tensor,
                     counts_in: Optional[torch.Tensor] = None, counts_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = x[b_i:e_i].sum(0) for one or two row ranges in ONE launch (fixed summation order); counts_out[i] =
    counts_in[i] rides along.  The payload of the centre statistics all-reduce."""
    _chk_cuda(x, out)
    n = len(segments)
    assert n in (1, 2) and x.dim() == 2 and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n * x.shape[1]
    K, ld = x.shape[1], _rowmajor(x)
    r = [e - b for b, e in segments] + [0]
    ws = zeroed_workspace("segment_cols_sum", int(_ext.lib().dinox_segment_cols_sum_workspace_bytes(r[0], r[1], K)), x.device)
    begin = (ctypes.c_int64 * n)(*[int(b) for b, _ in segments])
    end = (ctypes.c_int64 * n)(*[int(e) for _, e in segments])
    _ext.call("dinox_segment_cols_sum", _p(x), DT[x.dtype], K, ld, n, begin, end, _p(out), _p(counts_in), _p(counts_out),
              _p(ws), _stream())
    return out


def gather_f32(src: torch.Tensor, idx: torch.Tensor, fill: float = 0.0) -> torch.Tensor:
    out = torch.empty(idx.numel(), dtype=torch.float32, device=src.device)
    _ext.call("dinox_gather_f32", _p(src), _p(idx), idx.numel(), float(fill), _p(out), _stream())
    return out


def scatter_rows(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[idx[r]] = src[r] (fp32 rows; unique idx; idx < 0 skipped).  dst rows not named by idx keep their value."""
    _chk_cuda(src, dst, idx)
    assert src.dim() == 2 and dst.dim() == 2 and src.dtype == dst.dtype == torch.float32 and idx.dtype == torch.int64
    assert src.stride(1) == 1 and dst.stride(1) == 1 and idx.numel() == src.shape[0] and src.shape[1] == dst.shape[1]
    _ext.call("dinox_scatter_rows_f32", _p(src), src.stride(0), _p(idx), src.shape[0], src.shape[1], _p(dst),
              dst.stride(0), _stream())
    return dst


def gather_rows_f32(src: torch.Tensor, idx: Optional[torch.Tensor], out: torch.Tensor) -> torch.Tensor:
    """out[r] = float(src[idx[r]]) (idx None: identity; idx < 0: zero row); src rows may be strided."""
    _chk_cuda(src, out)
    assert src.dim() == 2 and src.stride(1) == 1 and out.stride(1) == 1 and out.dtype == torch.float32
    _ext.call("dinox_gather_rows_f32", _p(src), DT[src.dtype], src.stride(0), _p(idx), out.shape[0], src.shape[1], _p(out),
              out.stride(0), _stream())
    return out


def scatter_add_rows(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[idx[r]] += src[r] (fp32 rows; unique idx; idx < 0 skipped), in place."""
    _chk_cuda(src, dst, idx)
    assert src.dim() == 2 and dst.dim() == 2 and src.dtype == dst.dtype == torch.float32 and idx.dtype == torch.int64
    assert src.stride(1) == 1 and dst.stride(1) == 1 and idx.numel() == src.shape[0] and src.shape[1] == dst.shape[1]
    _ext.call("dinox_scatter_add_rows_f32", _p(src), src.stride(0), _p(idx), src.shape[0], src.shape[1], _p(dst),
              dst.stride(0), _stream())
    return dst


def gelu_fwd(a: torch.Tensor) -> torch.Tensor:
    h = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    _ext.call("dinox_gelu_fwd", _p(a), a.numel(), _p(h), _stream())
    return h


def gelu_bwd(dh: torch.Tensor, a: torch.Tensor, scale_dev: Optional[torch.Tensor] = None):
    rows, D = a.shape
    da = torch.empty(rows, D, dtype=torch.bfloat16, device=a.device)
    n_part = int(_ext.lib().dinox_gelu_bwd_workspace_bytes(rows, D)) // (4 * D)
    part = torch.empty(n_part, D, dtype=torch.float32, device=a.device)
    _ext.call("dinox_gelu_bwd", _p(dh), _p(a), rows, D, _p(scale_dev), _p(da), _p(part), _stream())
    return da, part


def gelu_bwd_gather(src: torch.Tensor, ptr: torch.Tensor, ent: torch.Tensor, a: torch.Tensor,
                    scale_dev: Optional[torch.Tensor] = None):
    """gelu_bwd on dh[r] = sum of the rows ent[ptr[r]:ptr[r+1]] of src ((E, D) fp32 or (S, E, D) split-K slabs, summed
    too) without materialising dh: gather_sum_rows + gelu_bwd in one launch, same summation order."""
    rows, D = a.shape
    if src.dim() == 3:
        slabs, slab_stride, ld = src.shape[0], src.stride(0), src.stride(1)
        assert src.stride(2) == 1
    else:
        slabs, slab_stride, ld = 1, 0, _rowmajor(src)
    da = torch.empty(rows, D, dtype=torch.bfloat16, device=a.device)
    n_part = int(_ext.lib().dinox_gelu_bwd_workspace_bytes(rows, D)) // (4 * D)
    part = torch.empty(n_part, D, dtype=torch.float32, device=a.device)
    _ext.call("dinox_gelu_bwd_gather", _p(src), ld, slabs, slab_stride, _p(ptr), _p(ent), _p(a), rows, D, _p(scale_dev),
              _p(da), _p(part), _stream())
    return da, part


def gemv_bf16(w: torch.Tensor, x: torch.Tensor, alpha: float = 1.0, bias: Optional[torch.Tensor] = None,
              beta: float = 1.0) -> torch.Tensor:
    K, D = w.shape
    out = torch.empty(K, dtype=torch.float32, device=w.device)
    _ext.call("dinox_gemv_bf16", _p(w), _rowmajor(w), _p(x), K, D, float(alpha), _p(bias), float(beta), _p(out), _stream())
    return out


def gemv_bf16_multi(w: torch.Tensor, xs: torch.Tensor, alphas: Sequence[float], bias: Optional[torch.Tensor] = None,
                    beta: float = 1.0, divisors: Optional[torch.Tensor] = None) -> torch.Tensor:
    """xs: (nvec, D) fp32 -> out (nvec, K): alphas[v] / divisors[v] * W @ xs[v] + beta * bias, one pass over W
    (divisors: optional (nvec,) fp32 DEVICE tensor)."""
    K, D = w.shape
    nvec = xs.shape[0]
    assert xs.is_contiguous() and xs.shape[1] == D and len(alphas) == nvec
    assert divisors is None or (divisors.dtype == torch.float32 and divisors.numel() == nvec and divisors.is_contiguous())
    out = torch.empty(nvec, K, dtype=torch.float32, device=w.device)
    al = (ctypes.c_float * nvec)(*[float(a) for a in alphas])
    _ext.call("dinox_gemv_bf16_multi", _p(w), _rowmajor(w), _p(xs), nvec, K, D, al, _p(divisors), _p(bias), float(beta),
              _p(out), _stream())
    return out


def gemv_bf16_multi_ema_(w: torch.Tensor, xs: torch.Tensor, alphas: Sequence[float], bias: Optional[torch.Tensor],
                         targets: Sequence[torch.Tensor], momenta: Sequence[float],
                         divisors: Optional[torch.Tensor] = None) -> None:
    """targets[v] <- m_v * targets[v] + (1 - m_v) * (alphas[v] / divisors[v] * W @ xs[v] + bias), in place, one pass
    over W: the centre / patch-centre update of the fused path."""
    K, D = w.shape
    nvec = xs.shape[0]
    assert xs.is_contiguous() and xs.shape[1] == D and len(alphas) == nvec == len(targets) == len(momenta)
    for t in targets:
        assert t.dtype == torch.float32 and t.numel() == K and t.is_contiguous()
    al = (ctypes.c_float * nvec)(*[float(a) for a in alphas])
    mo = (ctypes.c_float * nvec)(*[float(m) for m in momenta])
    tg = (ctypes.c_void_p * nvec)(*[t.data_ptr() for t in targets])
    _ext.call("dinox_gemv_bf16_multi_ema", _p(w), _rowmajor(w), _p(xs), nvec, K, D, al, _p(divisors), _p(bias), 1.0, tg, mo,
              _stream())


def sum_slabs(parts: torch.Tensor, out: torch.Tensor, accumulate: bool = False, scale: float = 1.0,
              scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (+)= scale * parts.sum(0) for contiguous (S, ...) fp32 split-K slabs (fixed order)."""
    assert parts.is_contiguous() and out.is_contiguous() and parts[0].numel() == out.numel()
    _ext.call("dinox_sum_slabs", _p(parts), parts.shape[0], parts.stride(0), out.numel(), _p(scale_dev), float(scale),
              _p(out), int(accumulate), _stream())
    return out


def gather_sum_rows(src: torch.Tensor, ptr: torch.Tensor, ent: torch.Tensor, rows: int, out: torch.Tensor,
                    scale: float = 1.0, scale_dev: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """src: (E, D) fp32, or (S, E, D) split-K partial slabs that are summed as well."""
    if src.dim() == 3:
        slabs, slab_stride, ld, D = src.shape[0], src.stride(0), src.stride(1), src.shape[2]
        assert src.stride(2) == 1
    else:
        slabs, slab_stride, ld, D = 1, 0, _rowmajor(src), src.shape[1]
    _ext.call("dinox_gather_sum_rows", _p(src), ld, slabs, slab_stride, _p(ptr), _p(ent), rows, D, _p(scale_dev),
              float(scale), _p(out), _rowmajor(out), int(accumulate), _stream())
    return out


def koleo_fwd(z: torch.Tensor, eps: float):
    """KoLeo forward on head outputs z (R, K) [fp32 | bf16 | fp16].  Returns (loss, saved) where
    saved = (inv_norm, nn, dist) feeds koleo_bwd."""
    _chk_cuda(z)
    ld = _rowmajor(z)
    R, K = z.shape
    dev = z.device
    ldb = (K + 7) // 8 * 8
    zb = torch.empty(R, ldb, dtype=torch.bfloat16, device=dev)
    inv = torch.empty(R, dtype=torch.float32, device=dev)
    _ext.call("dinox_koleo_rownorm", _p(z), DT[z.dtype], R, K, ld, _p(inv), _p(zb), ldb, _stream())
    zv = zb[:, :K]
    parts = gemm_bf16_splitk(zv, zv)                       # cosine ranking only: (S, R, R) partial Gram slabs
    gram = torch.empty(R, R, dtype=torch.float32, device=dev)
    sum_slabs(parts, gram)
    nc = int(_ext.lib().dinox_koleo_candidates())
    cand = torch.empty(R, nc, dtype=torch.int32, device=dev)
    d2 = torch.empty(R, nc, dtype=torch.float32, device=dev)
    nn = torch.empty(R, dtype=torch.int32, device=dev)
    dist = torch.empty(R, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    _ext.call("dinox_koleo_fwd", _p(z), DT[z.dtype], R, K, ld, _p(inv), _p(gram), R, float(eps), _p(cand), _p(d2), _p(nn),
              _p(dist), _p(loss), _stream())
    return loss, (inv, nn, dist)


def koleo_bwd(z: torch.Tensor, saved, eps: float, upstream: torch.Tensor) -> torch.Tensor:
    inv, nn, dist = saved
    R, K = z.shape
    dz = torch.empty(R, K, dtype=z.dtype, device=z.device)
    up = upstream.to(torch.float32).reshape(1).contiguous()
    _ext.call("dinox_koleo_bwd", _p(z), DT[z.dtype], R, K, _rowmajor(z), _p(inv), _p(nn), _p(dist), float(eps), _p(up),
              _p(dz), K, _stream())
    return dz


def scalar_combine(terms: Sequence[torch.Tensor], weights: Sequence[float], scale: float = 1.0,
                   out: Optional[torch.Tensor] = None, out_unscaled: Optional[torch.Tensor] = None) -> torch.Tensor:
    """0-dim fp32 tensor scale * sum_i weights[i] * terms[i] (device scalars, one launch, fixed order);
    `out_unscaled` (optional) receives the sum without `scale`."""
    n = len(terms)
    assert 1 <= n <= 8 and len(weights) == n
    for t in terms:
        assert t.dtype == torch.float32 and t.numel() == 1 and t.is_cuda
    out = torch.empty((), dtype=torch.float32, device=terms[0].device) if out is None else out
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in terms])
    w = (ctypes.c_float * n)(*[float(x) for x in weights])
    _ext.call("dinox_scalar_combine", ptrs, w, n, float(scale), _p(out), _p(out_unscaled), _stream())
    return out


def scalar_fanout(upstream: torch.Tensor, weights: Sequence[float], scale: float = 1.0) -> torch.Tensor:
    """(n,) fp32: upstream * scale * weights[i] - the per-term gradients of scalar_combine."""
    n = len(weights)
    up = upstream if (upstream.dtype == torch.float32 and upstream.is_contiguous()) else upstream.float().contiguous()
    out = torch.empty(n, dtype=torch.float32, device=up.device)
    w = (ctypes.c_float * n)(*[float(x) for x in weights])
    _ext.call("dinox_scalar_fanout", _p(up), w, n, float(scale), _p(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# fp32-faithful contraction mode: hi/lo bf16 splits and three-GEMM products
# ------------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor):
    """(hi, lo) bf16 with x ~= hi + lo to ~16 mantissa bits.  x: (rows, cols) or (batch, rows, cols) with unit stride
    in the last dimension (any float dtype).  The outputs have a 16-byte row pitch (the TMA maps of the GEMMs need it)
    and are returned as views of shape x.shape."""
    _chk_cuda(x)
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols)
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < cols):
        x2 = x2.contiguous()
    ldd = (cols + 7) // 8 * 8
    hi = torch.empty(x2.shape[0], ldd, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x2.shape[0], ldd, dtype=torch.bfloat16, device=x.device)
    _ext.call("dinox_split_bf16", _p(x2), DT[x2.dtype], x2.shape[0], cols, x2.stride(0) if x2.shape[0] > 1 else cols,
              _p(hi), _p(lo), ldd, _stream())
    lead = tuple(x.shape[:-1])
    view = lambda t: t.view(*lead, ldd)[..., :cols]
    return view(hi), view(lo)


def gemm3(a, b, *, a_mn_major=False, b_mn_major=False, out=None, accumulate=False, alpha=1.0, alpha_dev=None, bias_n=None):
    """fp32-grade C (+)= alpha * A @ B^T from hi/lo splits a = (a_hi, a_lo), b = (b_hi, b_lo): three tensor-core GEMMs
    (hi.hi + hi.lo + lo.hi) accumulated in fp32."""
    (ah, al), (bh, bl) = a, b
    kw = dict(a_mn_major=a_mn_major, b_mn_major=b_mn_major, alpha=alpha, alpha_dev=alpha_dev)
    out = gemm_bf16(ah, bh, out=out, accumulate=accumulate, bias_n=bias_n, **kw)
    gemm_bf16(ah, bl, out=out, accumulate=True, **kw)
    gemm_bf16(al, bh, out=out, accumulate=True, **kw)
    return out


def gemm3_batched(a, b, *, a_mn_major=False, b_mn_major=False, alpha=1.0, alpha_dev=None):
    (ah, al), (bh, bl) = a, b
    kw = dict(a_mn_major=a_mn_major, b_mn_major=b_mn_major, alpha=alpha, alpha_dev=alpha_dev)
    out = gemm_bf16_batched(ah, bh, **kw)
    gemm_bf16_batched(ah, bl, out=out, accumulate=True, **kw)
    gemm_bf16_batched(al, bh, out=out, accumulate=True, **kw)
    return out


def gelu_fwd_f32(a: torch.Tensor) -> torch.Tensor:
    h = torch.empty_like(a)
    _ext.call("dinox_gelu_fwd_f32", _p(a), a.numel(), _p(h), _stream())
    return h


def gelu_bwd_f32(dh: torch.Tensor, a: torch.Tensor, scale_dev: Optional[torch.Tensor] = None):
    rows, D = a.shape
    da = torch.empty(rows, D, dtype=torch.float32, device=a.device)
    n_part = int(_ext.lib().dinox_gelu_bwd_f32_workspace_bytes(rows, D)) // (4 * D)
    part = torch.empty(n_part, D, dtype=torch.float32, device=a.device)
    _ext.call("dinox_gelu_bwd_f32", _p(dh), _p(a), rows, D, _p(scale_dev), _p(da), _p(part), _stream())
    return da, part


def normalize_tokens_f32(feats: torch.Tensor, skip: int = 1):
    _chk_cuda(feats)
    Bt, T, D = feats.shape
    if feats.stride(2) != 1:
        feats = feats.contiguous()
    xn = torch.empty(Bt, T - skip, D, dtype=torch.float32, device=feats.device)
    inv = torch.empty(Bt, T - skip, dtype=torch.float32, device=feats.device)
    _ext.call("dinox_normalize_tokens_f32", _p(feats), DT[feats.dtype], Bt, T, D, feats.stride(0), feats.stride(1), skip,
              _p(xn), _p(inv), _stream())
    return xn, inv


def sqdiff(a: torch.Tensor, b: torch.Tensor, scale: float, want_delta: bool = True):
    """(scale * sum (a - b)^2, a - b) for contiguous fp32 tensors of one shape."""
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype == torch.float32
    delta = torch.empty_like(a) if want_delta else None
    loss = torch.empty((), dtype=torch.float32, device=a.device)
    ws = torch.empty(int(_ext.lib().dinox_sqdiff_workspace_bytes()), dtype=torch.uint8, device=a.device)
    _ext.call("dinox_sqdiff_f32", _p(a), _p(b), a.numel(), float(scale), _p(delta), _p(loss), _p(ws), _stream())
    return loss, delta


def head_offsets(b2s: torch.Tensor, b2t: torch.Tensor, center: torch.Tensor, center_patch: Optional[torch.Tensor],
                 inv_tau_s: float, inv_tau_t: float):
    """(cs2, ct2, ct2_patch | None): per-prototype log2-unit offsets of the fused passes, one launch."""
    K = b2s.numel()
    buf = torch.empty(3 if center_patch is not None else 2, K, dtype=torch.float32, device=b2s.device)
    _ext.call("dinox_head_offsets", _p(b2s), _p(b2t), _p(center), _p(center_patch), float(inv_tau_s), float(inv_tau_t),
              _p(buf[0]), _p(buf[1]), _p(buf[2]) if center_patch is not None else None, K, _stream())
    return buf[0], buf[1], (buf[2] if center_patch is not None else None)


def entry_weights(base: torch.Tensor, mask_weights: torch.Tensor, offset: int, scale: float) -> torch.Tensor:
    out = torch.empty_like(base)
    _ext.call("dinox_entry_weights", _p(base), base.numel(), _p(mask_weights), int(offset), mask_weights.numel(), float(scale),
              _p(out), _stream())
    return out


def fill_(t: torch.Tensor, v: float = 0.0) -> torch.Tensor:
    assert t.dtype == torch.float32 and t.is_contiguous()
    _ext.call("dinox_fill_f32", _p(t), t.numel(), float(v), _stream())
    return t


# ------------------------------------------------------------------------------------------------
# optional per-kernel timing with CUDA events on the launching stream (bench.py roofline leg)
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    """`with ops.TIMER.region("head_grad"):` brackets a launch with two CUDA events on the current
    stream when enabled; elapsed times are resolved after the caller synchronises."""

    def __init__(self):
        self.enabled = False
        self._pending = []
        self.totals = {}
        self.counts = {}
        self.samples = {}

    class _Region:
        def __init__(self, timer, name):
            self.t, self.name = timer, name

        def __enter__(self):
            if self.t.enabled:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            return self

        def __exit__(self, *exc):
            if self.t.enabled:
                self.e1.record()
                self.t._pending.append((self.name, self.e0, self.e1))
            return False

    def region(self, name):
        return KernelTimer._Region(self, name)

    def resolve(self):
        for name, e0, e1 in self._pending:
            self.totals[name] = self.totals.get(name, 0.0) + e0.elapsed_time(e1)
            self.counts[name] = self.counts.get(name, 0) + 1
            self.samples.setdefault(name, []).append(e0.elapsed_time(e1))
        self._pending = []

    def reset(self):
        self._pending, self.totals, self.counts, self.samples = [], {}, {}, {}


TIMER = KernelTimer()

"""dinox_b200 - B200-native (sm_100a) loss head for DINO-X.

Drop-in replacements for the reference's loss-head API (scripts/phase5_big_run.py:679-739,
:1798-1802; zoo/arch.py:246-261) backed by hand-written CUDA kernels behind a C ABI
(include/dinox_b200.h).  There is no CPU fallback: importing the compute API without the
built ``libdinox_b200.so`` raises.
"""
__version__ = "0.1.0"

_LAZY = {
    "DINOLoss", "DinoStudentTeacher", "compute_gram_matrix", "compute_gram_anchoring_loss",
    "ema_update", "_ema_update", "fused_head_dino_loss", "LossHead", "entropy_diagnostics",
    "sinkhorn_knopp_teacher", "KoLeoLoss", "FusedLossHead", "ProjectionHead", "combine_losses",
    "set_weight_cache", "invalidate_weight_cache", "set_contraction_precision", "token_fork",
}


def __getattr__(name):
    if name in ("FusedAdamW", "ShardedFusedAdamW"):
        from . import optim
        return getattr(optim, name)
    if name in _LAZY:
        from . import losshead
        return getattr(losshead, name)
    raise AttributeError(name)

"""Synthetic inputs for the loss-head path.

Two tiers:

* parity tier  - CT-like 2.5D crops + spacing vectors built the way the reference loader
  builds them (HU field -> 3 adjacent slices -> random window -> clip -> ImageNet normalise;
  scripts/phase5_big_run.py:246-249, :496, :520-523, :548-556; HU phantom in the spirit of
  scripts/preprocessing/phase2_preprocess_lidc_idri.py:197-205).  Pure CPU/torch, seeded.
* throughput tier - features drawn directly (cls ~ N(0,1), tokens ~ N(0,1)), head weights with
  PyTorch's default nn.Linear init (zoo/arch.py:252-256 builds the head outside
  PatchViT._init_weights, so xavier does not apply to it).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def seeded_generator(cfg: int, rank: int = 0) -> torch.Generator:
    """g = manual_seed(1000*cfg + rank): per-rank stream so that data-parallel shards differ."""
    g = torch.Generator()
    g.manual_seed(1000 * int(cfg) + int(rank))
    return g


def _hu_phantom(n_img: int, size: int, g: torch.Generator) -> torch.Tensor:
    """(n_img, 3, size, size) HU values: left-right gradient, one Gaussian lesion per image,
    a small per-slice offset (adjacent z) and N(0, 30 HU) noise."""
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32),
                            torch.arange(size, dtype=torch.float32), indexing="ij")
    base = xx / max(size - 1, 1) * 800.0 - 600.0
    cx = torch.rand(n_img, generator=g) * 0.6 + 0.2
    cy = torch.rand(n_img, generator=g) * 0.6 + 0.2
    sig = (torch.rand(n_img, generator=g) * 0.08 + 0.04) * size
    amp = torch.rand(n_img, generator=g) * 400.0 + 200.0
    blob = amp[:, None, None] * torch.exp(
        -((xx[None] - cx[:, None, None] * size) ** 2 + (yy[None] - cy[:, None, None] * size) ** 2)
        / (2.0 * sig[:, None, None] ** 2))
    z = torch.tensor([-1.0, 0.0, 1.0]) * 6.25  # +-50 HU over 16 slices -> 6.25 HU per slice
    vol = base[None, None] + blob[:, None] + z[None, :, None, None]
    vol = vol + torch.randn(vol.shape, generator=g) * 30.0
    return vol


def ct_crops(n_img: int, size: int, g: torch.Generator) -> torch.Tensor:
    """One windowed, normalised view per image: (n_img, 3, size, size), values in [-2.12, 2.64]."""
    hu = _hu_phantom(n_img, size, g)
    level = torch.rand(n_img, generator=g) * 800.0 - 400.0
    width = torch.rand(n_img, generator=g) * 1200.0 + 800.0
    wmin = (level - width / 2.0)[:, None, None, None]
    x = ((hu - wmin) / width[:, None, None, None]).clamp_(0.0, 1.0)
    mean = torch.tensor(IMAGENET_MEAN)[None, :, None, None]
    std = torch.tensor(IMAGENET_STD)[None, :, None, None]
    return (x - mean) / std


def spacing_vectors(n_img: int, g: torch.Generator) -> torch.Tensor:
    """(n_img, 3) mm spacings: sx=sy in U[0.46, 0.98], sz in U[0.625, 5.0] (LIDC-IDRI ranges)."""
    sxy = torch.rand(n_img, generator=g) * (0.98 - 0.46) + 0.46
    sz = torch.rand(n_img, generator=g) * (5.0 - 0.625) + 0.625
    return torch.stack([sxy, sxy, sz], dim=1)


def multicrop_batch(batch: int, g: torch.Generator, n_global: int = 2, n_local: int = 8,
                    global_size: int = 224, local_size: int = 96):
    """views: list of n_global (B,3,G,G) + n_local (B,3,L,L) tensors; spacing (B,3)."""
    views = [ct_crops(batch, global_size, g) for _ in range(n_global)]
    views += [ct_crops(batch, local_size, g) for _ in range(n_local)]
    return views, spacing_vectors(batch, g)


# ---------------------------------------------------------------------------------------------
# throughput tier
# ---------------------------------------------------------------------------------------------
@dataclass
class LossHeadShapes:
    """Row bookkeeping of one micro-step on one rank (SURVEY.md 8 notation)."""
    batch: int            # B images per rank
    dim: int              # D
    out_dim: int          # K
    n_patches: int        # N per global crop
    n_global: int = 2
    n_local: int = 8
    n_registers: int = 4
    mask_ratio: float = 0.3

    @property
    def views(self) -> int:
        return self.n_global + self.n_local

    @property
    def student_rows(self) -> int:      # Ms
        return self.batch * self.views

    @property
    def teacher_rows(self) -> int:      # Mt
        return self.batch * self.n_global

    @property
    def masked_per_crop(self) -> int:
        return int(math.floor(self.mask_ratio * self.n_patches))

    @property
    def masked_rows(self) -> int:       # Mm
        return self.teacher_rows * self.masked_per_crop

    @property
    def tokens(self) -> int:            # T = 1 + N + R
        return 1 + self.n_patches + self.n_registers

    def flops(self) -> float:
        """Algorithmic FLOPs per micro-step, SURVEY.md 8(d) / BASELINE.md 4."""
        D, K = self.dim, self.out_dim
        rows = 3 * self.student_rows + self.teacher_rows + 4 * self.masked_rows
        gram = 3 * self.teacher_rows * 2 * (self.n_patches + self.n_registers) ** 2 * D
        return 2.0 * D * D * rows + 2.0 * D * K * rows + gram

    def hbm_bytes(self, n_params: int, accum: int, sinkhorn: bool = False) -> float:
        """Algorithmic HBM bytes per micro-step (fused ideal), SURVEY.md 8(d)."""
        D, K = self.dim, self.out_dim
        Ms, Mt, Mm = self.student_rows, self.teacher_rows, self.masked_rows
        y = 2 * D * (Ms + Mt + 2 * Mm) + 4 * D * (Ms + Mm) + 6 * (K * D + D * D)
        y += 8 * (K * D + D * D + K + D) + 12 * (self.n_patches + self.n_registers) * D * Mt + 12 * K
        if sinkhorn:
            y += 28 * Mt * K
        y += 12.0 * n_params / accum
        return float(y)


def default_linear_init(out_f: int, in_f: int, g: torch.Generator) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for W and b."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2.0 - 1.0) * bound
    b = (torch.rand(out_f, generator=g) * 2.0 - 1.0) * bound
    return w, b


def head_weights(dim: int, out_dim: int, g: torch.Generator) -> Dict[str, torch.Tensor]:
    """State-dict of the projection head with the reference's keys (zoo/arch.py:252-256)."""
    w1, b1 = default_linear_init(dim, dim, g)
    w2, b2 = default_linear_init(out_dim, dim, g)
    return {"0.weight": w1, "0.bias": b1, "2.weight": w2, "2.bias": b2}


def masked_positions(shapes: LossHeadShapes, g: torch.Generator) -> torch.Tensor:
    """(Mt, n_masked) int64 patch indices: first floor(r*N) entries of a seeded randperm(N)
    per global crop."""
    n = shapes.masked_per_crop
    idx = [torch.randperm(shapes.n_patches, generator=g)[:n] for _ in range(shapes.teacher_rows)]
    return torch.stack(idx, dim=0) if n > 0 else torch.zeros(shapes.teacher_rows, 0, dtype=torch.long)


def feature_batch(shapes: LossHeadShapes, g: torch.Generator, with_tokens: bool = True,
                  with_ibot: bool = True, patches_from_tokens: bool = False) -> Dict[str, torch.Tensor]:
    """Backbone outputs for one micro-step, fp32 on CPU (caller moves / casts them).
    patches_from_tokens: the iBOT rows are not materialised; "patch_index" names them inside the (Mt, T, D)
    token tensors (flat row = crop * T + 1 + masked position: CLS first, then patches, then registers)."""
    out: Dict[str, torch.Tensor] = {}
    D = shapes.dim
    out["student_cls"] = torch.randn(shapes.student_rows, D, generator=g)
    out["teacher_cls"] = torch.randn(shapes.teacher_rows, D, generator=g)
    if with_tokens:
        out["student_tok"] = torch.randn(shapes.teacher_rows, shapes.tokens, D, generator=g)
        out["teacher_tok"] = torch.randn(shapes.teacher_rows, shapes.tokens, D, generator=g)
    if with_ibot and shapes.masked_rows > 0 and patches_from_tokens:
        assert with_tokens
        pos = masked_positions(shapes, g)                                  # (Mt, n_masked)
        crop = torch.arange(shapes.teacher_rows).unsqueeze(1)
        out["patch_index"] = (crop * shapes.tokens + 1 + pos).reshape(-1).to(torch.int64)
        out["masks_weight"] = torch.full((shapes.masked_rows,), 1.0 / shapes.masked_per_crop)
    elif with_ibot and shapes.masked_rows > 0:
        out["student_patch"] = torch.randn(shapes.masked_rows, D, generator=g)
        out["teacher_patch"] = torch.randn(shapes.masked_rows, D, generator=g)
        out["masks_weight"] = torch.full((shapes.masked_rows,), 1.0 / shapes.masked_per_crop)
    return out


CONFIGS = {
    # BASELINE.json configs (ViT-S/16: D=384, N=196; ViT-L/16: D=1024)
    "C1": dict(batch=8, dim=384, out_dim=65536, n_patches=196),
    "C2": dict(batch=64, dim=384, out_dim=65536, n_patches=196),
    "C3": dict(batch=32, dim=384, out_dim=65536, n_patches=196),
    "C4": dict(batch=32, dim=1024, out_dim=65536, n_patches=196),
    "C5lo": dict(batch=64, dim=384, out_dim=8192, n_patches=196),
    "C5": dict(batch=64, dim=384, out_dim=65536, n_patches=196),      # centre of the C5 sweep grid (bench.py --sweep)
    "C5hi": dict(batch=64, dim=384, out_dim=262144, n_patches=1024),  # top corner: K = 262144, 512x512 crops at patch 16
}
# BASELINE.json configs[4]: the C5 sweep K in [8192, 262144] x N in [196, 1024]
C5_SWEEP_K = (8192, 65536, 262144)
C5_SWEEP_N = (196, 576, 1024)


def student_param_shapes(dim: int, depth: int, out_dim: int, patch: int = 16, img: int = 224,
                         mlp_ratio: int = 4, scale_aware: bool = True) -> List[Tuple[int, ...]]:
    """Shapes of `DinoStudentTeacher(PatchViT(...), out_dim).parameters()` in iteration order
    (zoo/arch.py:150-261) - what the EMA loop walks."""
    n_patches = (img // patch) ** 2
    shapes: List[Tuple[int, ...]] = [(1, 1, dim), (1, 1 + n_patches, dim), (1, 4, dim), (dim, 3, patch, patch), (dim,)]
    if scale_aware:
        shapes += [(dim // 4, 3), (dim // 4,), (dim, dim // 4), (dim,), (dim,), (dim,)]
    for _ in range(depth):
        shapes += [(dim,), (dim,), (3 * dim, dim), (3 * dim,), (dim, dim), (dim,), (dim,), (dim,),
                   (mlp_ratio * dim, dim), (mlp_ratio * dim,), (dim, mlp_ratio * dim), (dim,)]
    shapes += [(dim,), (dim,), (dim, dim), (dim,), (out_dim, dim), (out_dim,)]
    return shapes


BACKBONES = {384: dict(depth=12), 1024: dict(depth=24)}  # ViT-S/16, ViT-L/16

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinox_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(2)
for (M, N, K) in [(640, 65536, 384), (640, 65536, 384), (512, 65536, 384), (1280, 32768, 384), (640, 65536, 64), (8064, 65536, 384)]:
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    b = torch.randn(N, K, generator=g).to(torch.bfloat16).to(dev)
    ref = a.float() @ b.float().t()
    for rep in range(2):
        out = ops.gemm_bf16(a, b)
        torch.cuda.synchronize()
        bad = ((out - ref).abs() > 1e-2 * ref.abs().max()).nonzero()
        if bad.shape[0] == 0:
            print(f"M={M} N={N} K={K} rep{rep}: OK")
        else:
            rows = bad[:, 0]; cols = bad[:, 1]
            mt = torch.unique(rows // 128).tolist(); ch = torch.unique(cols % 256 // 32).tolist(); nt = torch.unique(cols // 256).tolist()
            print(f"M={M} N={N} K={K} rep{rep}: {bad.shape[0]} bad; m_tiles {mt} chunk-in-tile {ch} n_tiles {nt[:12]} rows%128 range {int((rows%128).min())}-{int((rows%128).max())}")

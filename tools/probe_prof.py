"""Launch the four big GEMM flavours at C2 shapes a few times (for ncu captures)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinox_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(4)
E, K, D, rows = 8576, 65536, 384, 8064
hs = torch.randn(E, D, generator=g).to(torch.bfloat16).to(dev)
ht = torch.randn(E, D, generator=g).to(torch.bfloat16).to(dev)
ws = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
wt = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
cs2 = torch.zeros(K, device=dev); ct2 = torch.zeros(K, device=dev)
cw = torch.full((E,), 1.0 / E, device=dev)
loss = torch.zeros(2, device=dev)
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w2grad = torch.zeros(K, D, device=dev)
for _ in range(nrep):
    nat, l2 = ops.head_stats(hs[:rows], ws, 10.0, cs2)
    _, r2 = ops.head_stats(ht, wt, 25.0, ct2, want_nat=False)
    gt, db2p = ops.head_grad(ws, wt, hs, ht, 10.0, 25.0, cs2, ct2, None, 0, l2.new_zeros(E) + l2.mean(), r2, cw, loss)
    ops.gemm_bf16(gt, hs, b_mn_major=True, out=w2grad, accumulate=True, m_fastest=False)
    dh = ops.gemm_bf16_splitk(gt, ws, a_mn_major=True, b_mn_major=True)
torch.cuda.synchronize()
print("ok", loss.tolist())

"""Times of the big GEMM flavours at C2 shapes (CUDA events, 5 reps after 2 warm-ups)."""
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
w2grad = torch.zeros(K, D, device=dev)
nat, l2 = ops.head_stats(hs[:rows], ws, 10.0, cs2)
_, r2 = ops.head_stats(ht, wt, 25.0, ct2, want_nat=False)
lse_e = l2.new_zeros(E) + l2.mean()
gt = torch.empty(K, E, dtype=torch.bfloat16, device=dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = {}
t["stats"] = timeit(lambda: ops.head_stats(hs[:rows], ws, 10.0, cs2))
t["grad"] = timeit(lambda: ops.head_grad(ws, wt, hs, ht, 10.0, 25.0, cs2, ct2, None, 0, lse_e, r2, cw, loss, gt=gt))
t["dW2"] = timeit(lambda: ops.gemm_bf16(gt, hs, b_mn_major=True, out=w2grad, accumulate=True, m_fastest=False))
t["dH11"] = timeit(lambda: ops.gemm_bf16_splitk(gt, ws, a_mn_major=True, b_mn_major=True, splits=11))
t["dH2"] = timeit(lambda: ops.gemm_bf16_splitk(gt, ws, a_mn_major=True, b_mn_major=True, splits=2))
out = torch.empty(rows, K, dtype=torch.bfloat16, device=dev)
t["logits"] = timeit(lambda: ops.gemm_bf16(hs[:rows], ws, out=out))
print(os.environ.get("DINOX_LIB_TAG", "default"), "pair", os.environ.get("DINOX_PAIR", "0"),
      " ".join(f"{k} {v:.3f}" for k, v in t.items()), "loss", [round(x, 4) for x in loss.tolist()])

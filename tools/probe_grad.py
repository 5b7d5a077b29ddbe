"""head_grad / head_stats timing only (component-isolation builds give wrong numerics by design)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinox_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(4)
E, K, D, rows = 8576, 65536, int(os.environ.get("PROBE_D", "384")), 8064
hs = torch.randn(E, D, generator=g).to(torch.bfloat16).to(dev)
ht = torch.randn(E, D, generator=g).to(torch.bfloat16).to(dev)
ws = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
wt = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
cs2 = torch.zeros(K, device=dev); ct2 = torch.zeros(K, device=dev)
cw = torch.full((E,), 1.0 / E, device=dev)
loss = torch.zeros(2, device=dev)
lse_e = torch.full((E,), 12.0, device=dev); r2 = torch.full((E,), 14.0, device=dev)
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
print("D", D, os.environ.get("DINOX_LIB_TAG", "default"), "resa", os.environ.get("DINOX_RESA", "-"), "pair", os.environ.get("DINOX_PAIR", "-"),
      " ".join(f"{k} {v:.3f}" for k, v in t.items()))

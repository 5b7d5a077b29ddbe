"""One launch of every HBM-bound reduction kernel of the materialised-logit path at C2 shapes (target of an ncu
--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum capture; before every launch 512 MB are
rewritten and 256 MB read back, so the L2 is cold and holds no dirty lines)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinox_b200 import losshead, ops, synth
dev = "cuda"
sh = synth.LossHeadShapes(**synth.CONFIGS["C2"])
K, Ms, Mt, B = sh.out_dim, sh.student_rows, sh.teacher_rows, sh.batch
g = torch.Generator().manual_seed(7)
s = torch.randn(Ms, K, generator=g).to(dev); t = torch.randn(Mt, K, generator=g).to(dev)
center = torch.zeros(K, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros((), dtype=torch.int64, device=dev)
colb = ops.axpb(center, 25.0); rowb = ops.rows_lse(t, 25.0, colb); lse_s = ops.rows_lse(s, 10.0)
V, Vg = sh.views, sh.n_global
norm = 1.0 / ((Vg * V - Vg) * B)
up = torch.ones((), device=dev); colsum = ops.cols_sum(t)
n_params = 47085000
ps = [torch.randn(n_params // 8, device=dev) for _ in range(8)]; pt = [torch.randn_like(p) for p in ps]
plan = ops.EmaPlan(ps, pt)
for fn in (lambda: ops.rows_lse(s, 10.0), lambda: ops.rows_lse(t, 25.0, colb), lambda: ops.cols_lse(t, 25.0, rowb),
           lambda: ops.cols_sum(t), lambda: ops.ce_fwd(s, t, B, V, Vg, 10.0, 25.0, colb, rowb, lse_s, None, norm, True),
           lambda: ops.ce_fwd_onepass(s, t, B, V, Vg, 10.0, 25.0, colb, None, norm, True),
           lambda: ops.ce_bwd(s, t, B, V, Vg, 10.0, 25.0, colb, rowb, lse_s, None, norm, True, up),
           lambda: ops.center_ema_(center, colsum, Mt, 0.9), lambda: losshead.sinkhorn_knopp_biases(t, 0.04, 3, None),
           lambda: plan.apply(0.996)):
    flush.fill_(1); sink.copy_(flush.view(torch.int64)[: (256 << 20) // 8].max())   # cold AND clean L2
    fn()
torch.cuda.synchronize()
print("ok")

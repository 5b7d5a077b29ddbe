import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import dinox_b200 as dx
from oracle import losshead_oracle as O
g = np.load("/root/repo/tests/golden/koleo.npz")
for name in ["clustered"]:
    x = torch.from_numpy(g[f"{name}_x"]).cuda().requires_grad_(True)
    loss = dx.KoLeoLoss()(x); loss.backward()
    ref = torch.from_numpy(g[f"{name}_grad"])
    print(name, "loss", loss.item(), float(g[f"{name}_loss"]))
    d = (x.grad.cpu() - ref)
    print(" grad rel", (d.norm() / ref.norm()).item(), "rows with error:", (d.norm(dim=1) / ref.norm(dim=1).clamp_min(1e-30) > 1e-3).nonzero().flatten().tolist())
    # nearest neighbours per oracle
    xc = torch.from_numpy(g[f"{name}_x"])
    xn = torch.nn.functional.normalize(xc, dim=-1)
    pd = torch.cdist(xn.double(), xn.double()) + torch.eye(xc.shape[0]).double() * 1e9
    dmin, j = pd.min(1)
    pdf = torch.cdist(xn, xn) + torch.eye(xc.shape[0]) * 1e9
    dminf, jf = pdf.min(1)
    print(" fp64 nn == fp32-cdist nn:", (j == jf).all().item(), "mismatch rows", (j != jf).nonzero().flatten().tolist())
    print(" second-best gap (fp64) for mismatched rows:", [(pd[r].topk(2, largest=False).values.tolist()) for r in (j != jf).nonzero().flatten().tolist()][:4])

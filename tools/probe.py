"""GPU bring-up probe: runs one group of raw-kernel checks and prints numeric diagnostics.
Usage: python tools/probe.py <group>   (groups: env ema rows gemm stats grad autocast)
Each group runs in its own process so that a device-side trap in one cannot poison the others."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dinox_b200 import _ext, ops

dev = "cuda"


def relerr(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), (a - b).abs().max().item()


def g_env():
    print(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    print("cpu_count", os.cpu_count())
    print("device_check", _ext.lib().dinox_device_check())
    p = torch.cuda.get_device_properties(0)
    print("sms", p.multi_processor_count, "mem", p.total_memory)


def g_ema():
    g = torch.Generator().manual_seed(0)
    shapes = [(96,), (384, 96), (1, 197, 384), (1536, 384), (65536, 384), (7,), (1023,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    pt = [torch.randn(s, generator=g) for s in shapes]
    ref = [t.clone() for t in pt]
    m = 0.996
    for t, s in zip(ref, ps):
        t.mul_(m).add_(s, alpha=1.0 - m)
    ds = [x.to(dev) for x in ps]
    dt = [x.to(dev) for x in pt]
    plan = ops.EmaPlan(ds, dt)
    plan.apply(m)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(dt, ref)):
        d = (a.cpu() - b).abs().max().item()
        ulp = (a.cpu().view(torch.int32) - b.view(torch.int32)).abs().max().item()
        print(f"ema tensor {i} shape {tuple(b.shape)} maxabs {d:.3e} max ulp {ulp}")
    # bandwidth
    n = 64 * 1024 * 1024
    a = torch.randn(n, device=dev); b = torch.randn(n, device=dev)
    plan = ops.EmaPlan([a], [b])
    for _ in range(3): plan.apply(m)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.apply(m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"ema 64Mi elements: {ms:.3f} ms -> {12*n/ms/1e6:.1f} GB/s")


def g_rows():
    from oracle import losshead_oracle as O
    g = torch.Generator().manual_seed(1)
    for (B, V, Vg, K, dt) in [(4, 2, 2, 128, torch.float32), (3, 5, 2, 1000, torch.float32), (5, 2, 2, 4099, torch.float32),
                              (8, 10, 2, 8192, torch.float32), (8, 10, 2, 8192, torch.bfloat16), (4, 2, 2, 65536, torch.float32)]:
        s = torch.randn(B * V, K, generator=g).to(dt)
        t = torch.randn(B * Vg, K, generator=g).to(dt)
        c = torch.randn(1, K, generator=g) * 0.1
        ts, tt = 0.1, 0.04
        sg = s.float().clone().requires_grad_(True)
        loss = O.multicrop_dino_loss(sg, t.float(), c, ts, tt, Vg, V - Vg)
        loss.backward()
        sd, td, cd = s.to(dev), t.to(dev), c.to(dev).reshape(-1)
        colb = ops.axpb(cd, 1.0 / tt)
        rb, tent = ops.rows_lse(td, 1.0 / tt, colb, want_entropy=True)
        lse_s, sent = ops.rows_lse(sd, 1.0 / ts, None, want_entropy=True)
        n_terms = Vg * V - Vg
        l = ops.ce_fwd(sd, td, B, V, Vg, 1 / ts, 1 / tt, colb, rb, lse_s, None, 1.0 / (n_terms * B), True)
        up = torch.ones((), device=dev)
        gr = ops.ce_bwd(sd, td, B, V, Vg, 1 / ts, 1 / tt, colb, rb, lse_s, None, 1.0 / (n_terms * B), True, up)
        torch.cuda.synchronize()
        te, se = O.entropy_diagnostics(s.float(), t.float(), c, ts, tt)
        print(f"ce B={B} V={V} K={K} {dt}: loss {l.item():.7f} ref {loss.item():.7f} rel {abs(l.item()-loss.item())/abs(loss.item()):.2e} "
              f"grad rel/max {relerr(gr.cpu(), sg.grad)} ent t {tent.mean().item():.6f}/{te.item():.6f} s {sent.mean().item():.6f}/{se.item():.6f}")
        cs = ops.cols_sum(td)
        print("   cols_sum", relerr(cs.cpu(), t.float().sum(0)))
        center = cd.clone()
        ops.center_ema_(center, cs, t.shape[0], 0.9)
        print("   center", relerr(center.cpu(), O.center_update(c, t.float(), 0.9).reshape(-1)))
        cl = ops.cols_lse(td, 1.0 / tt, rb)
        ref = torch.logsumexp(t.float() / tt - rb.cpu()[:, None], dim=0)
        print("   cols_lse", relerr(cl.cpu(), ref))


def g_gemm():
    g = torch.Generator().manual_seed(2)
    cases = [
        (128, 256, 64, 0, 0), (128, 256, 384, 0, 0), (256, 512, 128, 0, 0), (200, 384, 384, 0, 0),
        (8064, 384, 384, 0, 0), (300, 1000, 200, 0, 0), (640, 65536, 384, 0, 0),
        (128, 256, 64, 0, 1), (256, 384, 128, 0, 1), (128, 256, 64, 1, 0), (128, 256, 64, 1, 1),
        (512, 384, 1024, 1, 1), (1000, 384, 520, 1, 1), (384, 200, 4096, 0, 1),
    ]
    for (M, N, K, am, bm) in cases:
        try:
            a = (torch.randn(M, K, generator=g) ).to(torch.bfloat16)
            b = (torch.randn(N, K, generator=g) ).to(torch.bfloat16)
            ref = a.float() @ b.float().t()
            ad = (a.t().contiguous() if am else a).to(dev)
            bd = (b.t().contiguous() if bm else b).to(dev)
            # pad leading dims to multiples of 8 elements for TMA
            def pad(t):
                if t.shape[1] % 8 == 0:
                    return t
                p = torch.zeros(t.shape[0], (t.shape[1] + 7) // 8 * 8, dtype=t.dtype, device=t.device)
                p[:, :t.shape[1]] = t
                return p[:, :t.shape[1]]
            ad, bd = pad(ad), pad(bd)
            out = ops.gemm_bf16(ad, bd, a_mn_major=bool(am), b_mn_major=bool(bm))
            torch.cuda.synchronize()
            r = relerr(out.cpu(), ref)
            print(f"gemm M={M} N={N} K={K} a_mn={am} b_mn={bm}: rel {r[0]:.3e} max {r[1]:.3e} {'OK' if r[0] < 1e-4 else 'MISMATCH'}")
            if r[0] >= 1e-4:
                d = (out.cpu() - ref)
                bad = (d.abs() > 1e-2 * ref.abs().max()).nonzero()
                print("   first bad idx", bad[:8].tolist(), "count", bad.shape[0], "of", d.numel())
                print("   out[0,:8]", out[0, :8].tolist()); print("   ref[0,:8]", ref[0, :8].tolist())
        except Exception as e:
            print(f"gemm M={M} N={N} K={K} a_mn={am} b_mn={bm}: EXC {e}")
            raise
    # epilogue options
    M, N, K = 256, 384, 128
    a = torch.randn(M, K, generator=g).to(torch.bfloat16); b = torch.randn(N, K, generator=g).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    ref = 0.5 * (a.float() @ b.float().t()) + bias
    c0 = torch.randn(M, N, generator=g)
    out = ops.gemm_bf16(a.to(dev), b.to(dev), alpha=0.5, bias_n=bias.to(dev))
    print("gemm alpha+bias", relerr(out.cpu(), ref))
    outb = ops.gemm_bf16(a.to(dev), b.to(dev), alpha=0.5, bias_n=bias.to(dev), out_dtype=torch.bfloat16)
    print("gemm bf16 out", relerr(outb.cpu(), ref))
    acc = c0.to(dev).clone()
    ops.gemm_bf16(a.to(dev), b.to(dev), out=acc, accumulate=True, alpha=2.0, alpha_dev=torch.tensor([0.25], device=dev))
    print("gemm accumulate", relerr(acc.cpu(), c0 + 0.5 * (a.float() @ b.float().t())))
    # ragged shapes through the TMA-store epilogue: fp32 / bf16 out, accumulate
    for (M, N, K) in [(200, 200, 384), (300, 1000, 200), (129, 72, 64)]:
        a = torch.randn(M, K, generator=g).to(torch.bfloat16); b = torch.randn(N, K, generator=g).to(torch.bfloat16)
        ref = a.float() @ b.float().t()
        ld = (N + 7) // 8 * 8
        buf = torch.full((M + 3, ld + 8), 7.0, device=dev)
        ops.gemm_bf16(a.to(dev), b.to(dev), out=buf[:M, :N])
        ok_guard = bool((buf[M:] == 7).all() and (buf[:, N:] == 7).all())
        print(f"gemm ragged f32 M={M} N={N} K={K}", relerr(buf[:M, :N].cpu(), ref), "guard", ok_guard)
        ops.gemm_bf16(a.to(dev), b.to(dev), out=buf[:M, :N], accumulate=True, alpha=-1.0)
        print(f"   accumulate -> ~0: max {buf[:M, :N].abs().max().item():.3e} guard", bool((buf[M:] == 7).all() and (buf[:, N:] == 7).all()))
        bb = torch.full((M + 3, ld + 8), 7.0, device=dev, dtype=torch.bfloat16)
        ops.gemm_bf16(a.to(dev), b.to(dev), out=bb[:M, :N])
        print(f"   bf16 out", relerr(bb[:M, :N].float().cpu(), ref), "guard", bool((bb[M:] == 7).all() and (bb[:, N:] == 7).all()))
    # batched (Gram shapes) and split-K
    Bt, T, D = 5, 200, 384
    x = torch.randn(Bt, T, D, generator=g).to(torch.bfloat16)
    ref = torch.bmm(x.float(), x.float().transpose(1, 2))
    out = ops.gemm_bf16_batched(x.to(dev), x.to(dev))
    print("gemm batched f32", relerr(out.cpu(), ref))
    outb = ops.gemm_bf16_batched(x.to(dev), x.to(dev), out_dtype=torch.bfloat16)
    print("gemm batched bf16", relerr(outb.float().cpu(), ref))
    for (M, N, K, S) in [(256, 384, 4096, 3), (200, 384, 8000, 7), (384, 384, 8064, None)]:
        a = torch.randn(K, M, generator=g).to(torch.bfloat16); b = torch.randn(K, N, generator=g).to(torch.bfloat16)
        ref = a.float().t() @ b.float()
        parts = ops.gemm_bf16_splitk(a.to(dev), b.to(dev), a_mn_major=True, b_mn_major=True, splits=S)
        print(f"gemm split-K M={M} N={N} K={K} splits={parts.shape[0]}", relerr(parts.sum(0).cpu(), ref))
    # the two backward GEMMs of the prototype layer at C2 shapes
    E, Kp, D = 8576, 65536, 384
    gt = (torch.randn(Kp, E, device=dev) * 0.01).to(torch.bfloat16)
    hs = torch.randn(E, D, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(Kp, D, device=dev) / math.sqrt(D)).to(torch.bfloat16)
    w2g = torch.zeros(Kp, D, device=dev)
    def timeit(fn, n=5):
        for _ in range(2): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms = timeit(lambda: ops.gemm_bf16(gt, hs, b_mn_major=True, out=w2g, accumulate=True, m_fastest=False))
    print(f"dW2 (K,E)x(E,D) accumulate: {ms:.3f} ms {2*E*Kp*D/ms/1e9:.1f} TF/s")
    ms = timeit(lambda: ops.gemm_bf16(gt, w2, a_mn_major=True, b_mn_major=True, m_fastest=True))
    print(f"dH unsplit: {ms:.3f} ms {2*E*Kp*D/ms/1e9:.1f} TF/s")
    for S in (2, 5, 11, 13):
        ms = timeit(lambda: ops.gemm_bf16_splitk(gt, w2, a_mn_major=True, b_mn_major=True, splits=S))
        print(f"dH split-K {S}: {ms:.3f} ms {2*E*Kp*D/ms/1e9:.1f} TF/s")
    ref = (gt[:, :256].float().t() @ w2.float())
    parts = ops.gemm_bf16_splitk(gt, w2, a_mn_major=True, b_mn_major=True)
    print("dH split-K parity (first 256 rows)", relerr(parts.sum(0)[:256].cpu(), ref.cpu()), "splits", parts.shape[0])
    del gt, hs, w2, w2g, parts
    # timing
    for (M, N, K) in [(8192, 8192, 8192), (8064, 65536, 384), (65536, 384, 8576)]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16); b = torch.randn(N, K, device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(2): ops.gemm_bf16(a, b, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.gemm_bf16(a, b, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        for _ in range(2): torch.matmul(a, b.t(), out=out)
        e0.record()
        for _ in range(5): torch.matmul(a, b.t(), out=out)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 5
        print(f"gemm time M={M} N={N} K={K}: ours {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TF/s | cublas {ms2:.3f} ms {2*M*N*K/ms2/1e9:.1f} TF/s")


def g_stats():
    g = torch.Generator().manual_seed(3)
    for (rows, K, D) in [(128, 256, 64), (200, 1000, 384), (640, 65536, 384), (8064, 65536, 384)]:
        h = torch.randn(rows, D, generator=g).to(torch.bfloat16)
        w = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16)
        col = torch.randn(K, generator=g) * 0.1
        inv_tau = 25.0
        hd, wd = h.to(dev), w.to(dev)
        col2 = (col * (inv_tau * ops.LOG2E)).to(dev)
        nat, l2 = ops.head_stats(hd, wd, inv_tau, col2)
        torch.cuda.synchronize()
        logits = (hd.float() @ wd.float().t()) * inv_tau + col.to(dev) * inv_tau
        ref = torch.logsumexp(logits, dim=-1)
        print(f"stats rows={rows} K={K} D={D}: {relerr(nat, ref)}  log2-consistency {relerr(l2 * math.log(2), nat)}")
        if rows >= 640:
            for _ in range(2): ops.head_stats(hd, wd, inv_tau, col2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ops.head_stats(hd, wd, inv_tau, col2)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"   time {ms:.3f} ms {2*rows*K*D/ms/1e9:.1f} TF/s")


def g_grad():
    g = torch.Generator().manual_seed(4)
    for (E, K, D) in [(128, 256, 64), (200, 1024, 384), (1152, 8192, 384), (8576, 65536, 384)]:
        hs = torch.randn(E, D, generator=g).to(torch.bfloat16).to(dev)
        ht = torch.randn(E, D, generator=g).to(torch.bfloat16).to(dev)
        ws = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
        wt = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
        b2s = (torch.randn(K, generator=g) * 0.05).to(dev); b2t = (torch.randn(K, generator=g) * 0.05).to(dev)
        cen = (torch.randn(K, generator=g) * 0.05).to(dev)
        cw = (torch.rand(E, generator=g) + 0.5).to(dev) / E
        its, itt = 10.0, 25.0
        S = (hs.float() @ ws.float().t() + b2s) * its
        T = (ht.float() @ wt.float().t() + b2t - cen) * itt
        lse_s = torch.logsumexp(S, -1); lse_t = torch.logsumexp(T, -1)
        p = torch.exp(S - lse_s[:, None]); q = torch.exp(T - lse_t[:, None])
        G = cw[:, None] * its * (p - q)
        loss_ref = (cw * (-(q * (S - lse_s[:, None])).sum(-1))).sum()
        cs2 = b2s * (its * ops.LOG2E); ct2 = (b2t - cen) * (itt * ops.LOG2E)
        loss = torch.zeros(2, device=dev)
        gt, db2p = ops.head_grad(ws, wt, hs, ht, its, itt, cs2, ct2, None, 0, lse_s * ops.LOG2E, lse_t * ops.LOG2E, cw, loss)
        torch.cuda.synchronize()
        print(f"grad E={E} K={K} D={D}: G {relerr(gt[:, :E].t(), G)} loss {loss[0].item():.6f} ref {loss_ref.item():.6f} "
              f"db2 {relerr(db2p.sum(0), G.sum(0))}")
        if E >= 1152:
            for _ in range(2): ops.head_grad(ws, wt, hs, ht, its, itt, cs2, ct2, None, 0, lse_s * ops.LOG2E, lse_t * ops.LOG2E, cw, loss, gt=gt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ops.head_grad(ws, wt, hs, ht, its, itt, cs2, ct2, None, 0, lse_s * ops.LOG2E, lse_t * ops.LOG2E, cw, loss, gt=gt)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"   time {ms:.3f} ms {2*2*E*K*D/ms/1e9:.1f} TF/s (2 GEMMs)")
        del S, T, p, q, G


def g_autocast():
    """SURVEY appendix B probe 1: dtypes under CUDA bf16 autocast of the ops the reference uses."""
    import torch.nn.functional as F
    x = torch.randn(8, 16, 64, device=dev)
    ln = torch.nn.LayerNorm(64).to(dev); lin = torch.nn.Linear(64, 128).to(dev)
    c = torch.zeros(1, 128, device=dev)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        f = ln(x); o = lin(f[:, 0]); ge = F.gelu(o)
        d = o - c; dd = d / 0.04; sm = F.softmax(dd, -1); st = o / 0.1; ls = F.log_softmax(st, -1)
        su = torch.sum(sm * ls, -1); mn = su.mean(); mc = torch.mean(o, 0, keepdim=True); ce = c * 0.9 + mc * 0.1
        nz = F.normalize(f[:, 1:], p=2, dim=-1); bm = torch.bmm(nz, nz.transpose(1, 2)); ms = F.mse_loss(bm, bm * 0.5)
    for n, t in [("layernorm", f), ("linear", o), ("gelu", ge), ("t-c", d), ("(t-c)/tau", dd), ("softmax", sm), ("s/tau", st),
                 ("log_softmax", ls), ("sum(t*s)", su), ("mean", mn), ("mean(t,0)", mc), ("center", ce), ("normalize", nz),
                 ("bmm", bm), ("mse", ms)]:
        print(f"autocast dtype {n}: {t.dtype}")


if __name__ == "__main__":
    grp = sys.argv[1]
    t0 = time.time()
    globals()["g_" + grp]()
    torch.cuda.synchronize()
    print(f"[{grp}] done in {time.time()-t0:.1f}s")

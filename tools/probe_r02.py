"""Round-2 probe: the read-back pass pair and the two backward GEMMs at C2 shapes.
  python tools/probe_r02.py time      CUDA-event medians of each launch (no profiler)
  python tools/probe_r02.py once      one launch of each (target of an ncu capture)"""
import math, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinox_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(4)
E, K, D, rows_s, rows_t = 8576, 65536, 384, 8064, 7552
if len(sys.argv) > 2:
    D = int(sys.argv[2])
hs = torch.randn(E, D, generator=g).to(torch.bfloat16).to(dev)
ht = torch.randn(rows_t, D, generator=g).to(torch.bfloat16).to(dev)
ws = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
wt = (torch.randn(K, D, generator=g) / math.sqrt(D)).to(torch.bfloat16).to(dev)
cs2 = torch.zeros(K, device=dev); ct2 = torch.zeros(K, device=dev)
cw = torch.full((E,), 1.0 / E, device=dev)
trow = (torch.arange(E, device=dev) % rows_t).to(torch.int32)
loss = torch.zeros(4, device=dev)
w2grad = torch.zeros(K, D, device=dev)
qt, refs = ops.teacher_buffers(rows_t, K, dev)
rb2_t = torch.empty(rows_t, device=dev)
G = torch.empty(E, K, dtype=torch.bfloat16, device=dev)
_, lse2 = ops.head_stats(hs, ws, 10.0, cs2, want_nat=False)


def teacher():
    ops.head_teacher(ht, wt, 25.0, ct2, None, 0, qt=qt, refs=refs, out_log2=rb2_t)


def stats():
    ops.head_stats(hs[:rows_s], ws, 10.0, cs2, want_nat=False)


def grad2(db2=True):
    rb2_e = rb2_t[trow.long()]
    return lambda: ops.head_grad2(hs, ws, 10.0, cs2, lse2, cw, rb2_e, trow, qt, refs, 1152, loss, want_db2=db2, g=G)


def dw2():
    ops.gemm_bf16(G, hs, a_mn_major=True, b_mn_major=True, out=w2grad, accumulate=True, m_fastest=False)


def dh():
    ops.gemm_bf16_splitk(G, ws, b_mn_major=True)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


teacher()
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
if mode == "once":
    stats(); teacher(); grad2()(); dw2(); dh()
    torch.cuda.synchronize()
    print("ok", loss.tolist())
else:
    res = {"head_stats_student": timeit(stats), "head_teacher": timeit(teacher), "head_grad2": timeit(grad2()),
           "head_grad2_nodb2": timeit(grad2(False)), "dW2": timeit(dw2), "dH": timeit(dh)}
    print({k: round(v, 4) for k, v in res.items()})

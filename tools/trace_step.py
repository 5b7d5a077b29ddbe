"""Timeline of one replayed micro-step from the library's launch trace (no nsys in this image).

    python tools/trace_step.py [--config C2] [--replays 20] [--out profiles/r02_step_timeline.txt]

dinox_trace_begin() makes every launch of the library be followed, on its stream, by a one-thread kernel that writes
%globaltimer.  Captured into the micro-step's CUDA graph, the stamps are rewritten by every replay: the END time of every
kernel, on every stream of the graph.  Per launch this prints the end time (us after the first stamp of the step), the
time since the previous stamp on the same stream (= the kernel's duration when it started right behind its
predecessor; an upper bound otherwise: it includes waiting for other streams / for SMs) and the stream.  The stamps
cost ~1-2 us each (both the traced and the untraced step time are printed).  Medians over the replays; every read-back
takes the LAST of six back-to-back replays (steady state) unless --isolated.
"""
import argparse
import ctypes
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dinox_b200 import _ext, synth
from dinox_b200.step import LossHeadStep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--replays", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--isolated", action="store_true", help="one replay per read-back (GPU idle in between) instead of the "
                    "last of six back-to-back replays")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    shapes = synth.LossHeadShapes(**synth.CONFIGS[a.config])
    feats = synth.feature_batch(shapes, synth.seeded_generator(2), patches_from_tokens=True)

    def make():
        st = LossHeadStep(shapes, dev, accum=1 << 30, with_backbone_params=False)     # no EMA inside the window
        slot = st.static_inputs(feats, slots=1)[0]
        with torch.no_grad():
            for k, v in feats.items():
                slot[k].copy_(v)
        return st

    def timed(st, n):
        for _ in range(3):
            st.micro_step_graph(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            st.micro_step_graph(0)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    plain = make()
    plain.capture(0)
    ms_plain = timed(plain, a.replays)

    cap = 1024
    slots = torch.zeros(cap, dtype=torch.int64, device=dev)
    traced = make()
    traced._ensure_grads()
    traced._fwd_bwd(traced._static[0])          # eager warm-up outside the trace (workspaces, plans)
    torch.cuda.synchronize()
    lib = _ext.lib()
    # capture() warms up twice before capturing: restart the trace right before the captured pass by hooking graph entry
    orig_graph = torch.cuda.graph

    class _G(orig_graph):
        def __enter__(self):
            r = super().__enter__()
            _ext.check(lib.dinox_trace_begin(ctypes.c_void_p(slots.data_ptr()), cap), "trace_begin")
            return r

        def __exit__(self, *exc):
            self_n[0] = lib.dinox_trace_end()
            return super().__exit__(*exc)
    self_n = [0]
    torch.cuda.graph = _G
    try:
        traced.capture(0)
    finally:
        torch.cuda.graph = orig_graph
    n = self_n[0]
    names = [lib.dinox_trace_name(i).decode() for i in range(n)]
    streams = [int(lib.dinox_trace_stream(i)) for i in range(n)]
    sid = {s: i for i, s in enumerate(dict.fromkeys(streams))}
    ms_traced = timed(traced, 5)
    ends = []
    for _ in range(a.replays):
        for _ in range(1 if a.isolated else 6):      # steady state: the stamps that survive are the LAST replay's
            traced.micro_step_graph(0)
        torch.cuda.synchronize()
        t = slots[:n].cpu().tolist()
        t0 = min(t)
        ends.append([(x - t0) / 1000.0 for x in t])
    med_end = [statistics.median(e[i] for e in ends) for i in range(n)]
    # time since the previous stamp on the same stream, per replay, then the median
    since = []
    for i in range(n):
        prev = [j for j in range(i) if streams[j] == streams[i]]
        if not prev:
            since.append(float("nan"))
            continue
        j = prev[-1]
        since.append(statistics.median(e[i] - e[j] for e in ends))
    order = sorted(range(n), key=lambda i: med_end[i])
    lines = [f"# {a.config}: one replayed micro-step, {n} launches on {len(sid)} streams; medians of {a.replays} replays",
             f"# untraced replay {ms_plain:.4f} ms per step, traced {ms_traced:.4f} ms (stamp kernels included)",
             "# end_us = end of the kernel after the first stamp of the step; since_us = time since the previous kernel END on the",
             "# same stream (duration if it started right away, else it includes waiting for another stream / for free SMs)",
             f"# {'end_us':>9} {'since_us':>9}  stream  kernel"]
    for i in order:
        lines.append(f"  {med_end[i]:9.1f} {since[i]:9.1f}  s{sid[streams[i]]}      {names[i]}")
    lines.append(f"# span first stamp -> last stamp: {max(med_end):.1f} us")
    txt = "\n".join(lines) + "\n"
    print(txt)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        open(a.out, "w").write(txt)


if __name__ == "__main__":
    main()

"""Read-only vs copy HBM bandwidth on this box with plain torch ops (best of 10, CUDA events): the denominator
MEASURED_PEAKS.json gives is COPY bandwidth (read + write bytes); this prints what a pure read stream reaches."""
import json, torch
dev = "cuda"
n = 1 << 28   # 1 GiB of fp32
a = torch.randn(n, device=dev); b = torch.empty_like(a)
def best(fn, nbytes, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return nbytes / (min(ts) * 1e-3) / 1e9
for _ in range(3):
    b.copy_(a); a.sum(); a.max()
out = {"copy_gbs_read_plus_write": best(lambda: b.copy_(a), 8.0 * n), "sum_gbs_read_only": best(lambda: a.sum(), 4.0 * n),
       "max_gbs_read_only": best(lambda: a.max(), 4.0 * n), "fill_gbs_write_only": best(lambda: b.fill_(1.0), 4.0 * n),
       "bytes": 4 * n}
print(json.dumps(out))

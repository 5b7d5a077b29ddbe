#!/usr/bin/env python
"""Loss-head throughput benchmark (BASELINE.json metric: loss-head crops/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPUs

A step = one micro-step of the loss head on one batch of synthetic backbone features: projection
head + multi-crop DINO CE + iBOT masked-patch CE (fused, logits never in HBM), Gram anchoring,
backward to d(features) and .grad of the head, centre updates, and the EMA of ALL student
parameters once per `accum` micro-steps.  Workload at N=1: BASELINE.json configs[1]
("C2": ViT-S/16, batch 64 x accum 4, 2 global + 8 local crops, K=65536, Gram on).  N>1: the same
per-GPU batch on every rank (weak scaling), centre statistics all-reduced over NCCL.

Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loss_head_crops_per_sec"
_OUT = sys.stdout


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.power, self.power_limit = [], None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:
                self.power_limit = None
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        pmax = max(self.power) if self.power else None
        reasons = set(self.reasons)
        # the power-cap flag is momentary: a board drawing its enforced limit while the SM clock sits
        # below max IS power capped even when no sample caught the flag
        if pmax is not None and self.power_limit and pmax >= 0.9 * self.power_limit and med and self.max_mhz and med < self.max_mhz:
            reasons.add("sw_power_cap")
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(self.samples),
                "power_w_max": pmax, "power_limit_w": self.power_limit}


# ---------------------------------------------------------------------------------------------------
# CPU legs: the oracle (a port of the reference algorithm) timed on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_oracle_step_time(cfg_name: str, sample_batch: int, steps: int, warmup: int, accum: int,
                         patches_from_tokens: bool = True):
    """Times oracle.LossHeadOracle.step (+ amortised EMA of the full parameter list) on a
    `sample_batch`-image slice of the workload.  Returns (crops/s, cores, description)."""
    from dinox_b200 import synth
    from oracle import losshead_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(synth.CONFIGS[cfg_name])
    cfg["batch"] = sample_batch
    sh = synth.LossHeadShapes(**cfg)
    g = synth.seeded_generator(2, 0)
    D, K = sh.dim, sh.out_dim
    sw, tw = synth.head_weights(D, K, g), synth.head_weights(D, K, g)
    sp = O.HeadParams(*[sw[k].requires_grad_(True) for k in ("0.weight", "0.bias", "2.weight", "2.bias")])
    tp = O.HeadParams(*[tw[k] for k in ("0.weight", "0.bias", "2.weight", "2.bias")])
    orc = O.LossHeadOracle(sp, tp, K, center_momentum=0.9, n_global=sh.n_global, n_local=sh.n_local, policy="fp32")
    depth = synth.BACKBONES.get(D, dict(depth=12))["depth"]
    shapes = synth.student_param_shapes(D, depth, K)[:-4]
    s_all = [torch.randn(s) * 0.02 for s in shapes] + [p.detach() for p in sp.tensors()]
    t_all = [torch.randn(s) * 0.02 for s in shapes] + [p.detach().clone() for p in tp.tensors()]
    f = synth.feature_batch(sh, g, patches_from_tokens=patches_from_tokens)
    times = []
    for i in range(warmup + steps):
        fs = {k: (v.clone().requires_grad_(True) if k.startswith("student") else v) for k, v in f.items()}
        t0 = time.perf_counter()
        if "patch_index" in fs:   # the iBOT rows are the masked rows of the token tensors (gathered inside the step)
            fs["student_patch"] = fs["student_tok"].reshape(-1, D)[fs["patch_index"]]
            fs["teacher_patch"] = fs["teacher_tok"].reshape(-1, D)[fs["patch_index"]]
        orc.step(fs["student_cls"], fs["teacher_cls"], 0.1, 0.04, student_tok=fs["student_tok"],
                 teacher_tok=fs["teacher_tok"], student_patch=fs.get("student_patch"),
                 teacher_patch=fs.get("teacher_patch"), masks_weight=fs.get("masks_weight"), accum=accum)
        if (i + 1) % accum == 0:
            orc.ema(s_all, t_all, 0.996)
            for p in sp.tensors():
                p.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    desc = (f"{cfg_name} shapes at batch {sample_batch} ({sh.student_rows} student / {sh.teacher_rows} teacher / "
            f"{sh.masked_rows} masked rows, K={K}), {steps} steps after {warmup} warm-up, fp32, "
            f"torch {torch.__version__} CPU, oracle port of the reference algorithm")
    return sh.student_rows / per_step, cores, desc, per_step


def torch_eager_gpu_step_time(cfg_name: str, dev, steps: int, warmup: int, accum: int):
    """The same-box comparator: the reference algorithm as plain PyTorch eager ops on THIS GPU
    under torch.autocast(bf16) - the oracle's code executed on CUDA, i.e. what scripts/phase5_big_run.py --amp would
    launch for this workload (cuBLAS GEMMs + ATen softmax / reductions, logits materialised) - with the reference's
    per-tensor EMA loop.  Full configuration, inputs resident.  Returns (crops/s, ms per micro-step)."""
    from dinox_b200 import synth
    from oracle import losshead_oracle as O
    sh = synth.LossHeadShapes(**synth.CONFIGS[cfg_name])
    g = synth.seeded_generator(2, 0)
    D, K = sh.dim, sh.out_dim
    sw, tw = synth.head_weights(D, K, g), synth.head_weights(D, K, g)
    keys = ("0.weight", "0.bias", "2.weight", "2.bias")
    sp = O.HeadParams(*[sw[k].to(dev).requires_grad_(True) for k in keys])
    tp = O.HeadParams(*[tw[k].to(dev) for k in keys])
    orc = O.LossHeadOracle(sp, tp, K, center_momentum=0.9, n_global=sh.n_global, n_local=sh.n_local, policy="fp32")
    orc.center, orc.center_patch = orc.center.to(dev), orc.center_patch.to(dev)
    depth = synth.BACKBONES.get(D, dict(depth=12))["depth"]
    shapes = synth.student_param_shapes(D, depth, K)[:-4]
    s_all = [torch.randn(s, device=dev) * 0.02 for s in shapes] + [p.detach() for p in sp.tensors()]
    t_all = [torch.randn(s, device=dev) * 0.02 for s in shapes] + [p.detach().clone() for p in tp.tensors()]
    f = {k: v.to(dev) for k, v in synth.feature_batch(sh, g, patches_from_tokens=True).items()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            e0.record()
        fs = {k: (v.clone().requires_grad_(True) if k.startswith("student") else v) for k, v in f.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            sp_rows = fs["student_tok"].reshape(-1, D)[fs["patch_index"]]
            tp_rows = fs["teacher_tok"].reshape(-1, D)[fs["patch_index"]]
            orc.step(fs["student_cls"], fs["teacher_cls"], 0.1, 0.04, student_tok=fs["student_tok"],
                     teacher_tok=fs["teacher_tok"], student_patch=sp_rows, teacher_patch=tp_rows,
                     masks_weight=fs["masks_weight"], accum=accum)
        if (i + 1) % accum == 0:
            orc.ema(s_all, t_all, 0.996)
            for p in sp.tensors():
                p.grad = None
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del orc, sp, tp, s_all, t_all, f
    torch.cuda.empty_cache()
    return sh.student_rows / (ms * 1e-3), ms


def reference_verbatim_cpu(batch: int, steps: int, warmup: int, K: int = 65536, D: int = 384, tokens: int = 201):
    """The reference's OWN classes (imported from /root/reference when that tree exists - it does in the build
    container, not on the GPU box) on the host cores: head forward for student and teacher, DINOLoss, Gram anchoring,
    backward, and the inline EMA loop of scripts/phase5_big_run.py:1798-1802 over the head parameters, on the
    2-global-view subset the reference can run (no local crops, no iBOT).  Returns None when the tree is absent."""
    ref_root = os.environ.get("DINOX_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "scripts")):
        return None
    import importlib
    for pth in (ref_root, os.path.join(ref_root, "scripts")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    try:
        big = importlib.import_module("phase5_big_run")
    except Exception as exc:   # missing optional dependency of the reference script
        return {"unavailable": f"{type(exc).__name__}: {exc}"}
    from dinox_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = synth.seeded_generator(2, 0)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(D, D), torch.nn.GELU(), torch.nn.Linear(D, K))   # zoo/arch.py:252-256
    s_head, t_head = mk(), mk()
    s_head.load_state_dict(synth.head_weights(D, K, g))
    t_head.load_state_dict(synth.head_weights(D, K, g))
    for p in t_head.parameters():
        p.requires_grad_(False)
    loss_fn = big.DINOLoss(K, 0.9)
    sf = torch.randn(2 * batch, tokens, D, generator=g)
    tf = torch.randn(2 * batch, tokens, D, generator=g)
    times = []
    for i in range(warmup + steps):
        x = sf.clone().requires_grad_(True)
        t0 = time.perf_counter()
        student_out = s_head(x[:, 0])
        with torch.no_grad():
            teacher_out = t_head(tf[:, 0])
        loss = loss_fn(student_out, teacher_out, 0.1, 0.04) + 1.0 * big.compute_gram_anchoring_loss(x, tf)
        loss.backward()
        with torch.no_grad():
            for p_s, p_t in zip(s_head.parameters(), t_head.parameters()):
                p_t.data.mul_(0.996).add_(p_s.data, alpha=1 - 0.996)
        for p in s_head.parameters():
            p.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per = statistics.median(times)
    return {"value": 2 * batch / per, "unit": "crops/s", "ms_per_step": per * 1e3, "cores": cores, "kind": "reference",
            "sample": f"reference classes verbatim (DINOLoss, compute_gram_anchoring_loss, nn.Sequential head, inline EMA of the head) "
                      f"on the 2-global-view subset: batch {batch} -> {2 * batch} crops, K={K}, D={D}, {tokens} tokens, fp32, "
                      f"median of {steps} steps after {warmup} warm-up"}


def workload_text(args, sh, n_params: float) -> str:
    """The workload both arms run (BASELINE.json configs), in the same words."""
    return (f"{args.config}: ViT-{'S' if sh.dim == 384 else 'L'}/16 loss head (teacher {args.teacher_mode}), per-GPU batch {sh.batch} x accum {args.accum}, "
            f"{sh.n_global} global + {sh.n_local} local crops, K={sh.out_dim}, D={sh.dim}, iBOT r={sh.mask_ratio} "
            f"({sh.masked_rows} masked rows, {'gathered from the token tensors by index' if args.patch_source == 'tokens' else 'materialised by the caller'}), "
            f"Gram anchoring on ({sh.tokens - 1} tokens), EMA of {n_params / 1e6:.1f} M params every {args.accum} micro-steps")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dinox_b200 import synth
    value, cores, desc, per_step = cpu_oracle_step_time(args.config, args.cpu_sample_batch, args.steps, args.warmup,
                                                         args.accum, args.patch_source == "tokens")
    sh = synth.LossHeadShapes(**synth.CONFIGS[args.config])
    depth = synth.BACKBONES.get(sh.dim, dict(depth=12))["depth"]
    n_params = 0
    for shp in synth.student_param_shapes(sh.dim, depth, sh.out_dim):
        n_params += math.prod(shp)
    verbatim = None
    if not args.no_reference_verbatim:
        v8 = reference_verbatim_cpu(args.cpu_sample_batch, 3, 1, K=sh.out_dim, D=sh.dim, tokens=sh.tokens)
        if v8 is not None and "value" in v8:
            verbatim = {"slice": v8, "full_batch": reference_verbatim_cpu(sh.batch, 2, 1, K=sh.out_dim, D=sh.dim, tokens=sh.tokens)}
        else:
            verbatim = v8 or {"unavailable": "the reference tree (/root/reference) is not present on this host: the reference "
                                             "is a Python package that cannot travel to the GPU box; see "
                                             "profiles/r02_cpu_reference_verbatim.json for the run in the build container"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args, sh, n_params),
                   "sample": f"each step runs a {args.cpu_sample_batch}-image slice of the per-GPU batch on {cores} host threads; "
                             f"crops/s = slice crops / slice time"},
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cores, "kind": "port", "sample": desc,
                         "processes": 1, "note": "ONE CPU process on this host's cores at every --gpus N"},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_verbatim": verbatim,
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------
# HBM-bound kernels of the materialised-logit path (DINOLoss drop-in, Sinkhorn-Knopp, centre EMA): achieved GB/s
# ---------------------------------------------------------------------------------------------------
def hbm_kernel_table(dev, sh, hbm_peak: float, reps: int = 20, only=None):
    """CUDA-event medians of the row / column reduction kernels at this config's logit shapes, each launched on
    a cold, CLEAN L2 (a 512 MB buffer is rewritten and half of it read back between calls), with their ALGORITHMIC bytes (every logit read once
    per pass, fp32) -> GB/s and fraction of the measured HBM copy bandwidth.  Shapes at C2: student 640 x 65536,
    teacher 128 x 65536 - the small ones are launch-latency bound and say so."""
    from dinox_b200 import losshead, ops
    K, Ms, Mt, B = sh.out_dim, sh.student_rows, sh.teacher_rows, sh.batch
    g = torch.Generator(device="cpu").manual_seed(7)
    s = torch.randn(Ms, K, generator=g).to(dev)
    t = torch.randn(Mt, K, generator=g).to(dev)
    center = torch.zeros(K, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    flush_words = flush.view(torch.int64)
    sink = torch.zeros((), dtype=torch.int64, device=dev)
    colb = ops.axpb(center, 25.0)
    rowb = ops.rows_lse(t, 25.0, colb)
    lse_s = ops.rows_lse(s, 10.0)
    V, Vg = sh.views, sh.n_global
    norm = 1.0 / ((Vg * V - Vg) * B)
    up = torch.ones((), device=dev)
    colsum = ops.cols_sum(t)
    cases = {
        "rows_lse_student": (lambda: ops.rows_lse(s, 10.0), 4.0 * Ms * K),
        "rows_lse_teacher": (lambda: ops.rows_lse(t, 25.0, colb), 4.0 * Mt * K + 4.0 * K),
        "cols_lse_teacher": (lambda: ops.cols_lse(t, 25.0, rowb), 4.0 * Mt * K + 4.0 * K),
        "cols_sum_teacher": (lambda: ops.cols_sum(t), 4.0 * Mt * K + 4.0 * K),
        "ce_fwd": (lambda: ops.ce_fwd(s, t, B, V, Vg, 10.0, 25.0, colb, rowb, lse_s, None, norm, True), 4.0 * (Ms + Mt) * K),
        "ce_fwd_onepass": (lambda: ops.ce_fwd_onepass(s, t, B, V, Vg, 10.0, 25.0, colb, None, norm, True),
                           4.0 * (Ms + Mt) * K),
        "ce_bwd": (lambda: ops.ce_bwd(s, t, B, V, Vg, 10.0, 25.0, colb, rowb, lse_s, None, norm, True, up),
                   4.0 * (2 * Ms + Mt) * K),
        "center_ema": (lambda: ops.center_ema_(center, colsum, Mt, 0.9), 12.0 * K),
        "sinkhorn_3it": (lambda: losshead.sinkhorn_knopp_biases(t, 0.04, 3, None), 6.0 * 4.0 * Mt * K),
    }
    out = {}
    for name, (fn, nbytes) in cases.items():
        if only and name not in only:
            continue
        fn()
        ts = []
        for _ in range(reps):
            # evict the logits from L2 (not timed): rewrite 512 MB, then READ 256 MB of it so that the cache is left
            # full of CLEAN lines - after the write alone up to 126 MB of dirty lines are written back during the
            # timed kernel (a 168 MB read-only pass then moves 294 MB of HBM traffic and reads as 0.5 of peak)
            flush.fill_(1)
            sink.copy_(flush_words[: (256 << 20) // 8].max())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(statistics.median(ts))
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
    return out


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from dinox_b200 import ops, synth
    from dinox_b200.step import LossHeadStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = True

    sh = synth.LossHeadShapes(**synth.CONFIGS[args.config])
    step = LossHeadStep(sh, dev, accum=args.accum, process_group=pg, teacher_mode=args.teacher_mode)
    g = synth.seeded_generator(2, rank)
    # iBOT rows: named by index inside the (crops, T, D) token tensors the backbone emits (default; the step
    # gathers them on the device) or shipped as separately materialised rows (--patch-source rows)
    feats = synth.feature_batch(sh, g, patches_from_tokens=(args.patch_source == "tokens"))

    def blob_views(blob, like):
        """typed views into one contiguous byte blob, 256-byte aligned segments, same keys/shapes as `like`"""
        views, off = {}, 0
        for k, v in like.items():
            n = v.numel() * v.element_size()
            views[k] = blob[off:off + n].view(v.dtype).view(v.shape)
            off += (n + 255) // 256 * 256
        return views

    blob_bytes = sum((v.numel() * v.element_size() + 255) // 256 * 256 for v in feats.values())
    host_blob = torch.empty(blob_bytes, dtype=torch.uint8).pin_memory()   # ONE pinned buffer -> ONE H2D copy per step
    host = blob_views(host_blob, feats)
    for k, v in feats.items():
        host[k].copy_(v)
    h2d_bytes = blob_bytes

    def to_leaves(f):
        return {k: (v.requires_grad_(True) if k.startswith("student") else v) for k, v in f.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def per_rank(ms: float):
        """[ms of every rank] (the same list on every rank)"""
        if world == 1:
            return [ms]
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def replay_leg(st, feats_like, steps, warmup):
        """graph-replay timing of another configuration (inputs resident): ms per micro-step, max over ranks"""
        slots = st.static_inputs(feats_like, slots=1)
        for k, v in feats_like.items():
            slots[0][k].detach().copy_(v)
        st.capture(0)
        for _ in range(warmup):
            st.micro_step_graph(0)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            st.micro_step_graph(0)
        b.record()
        barrier()
        mine = a.elapsed_time(b) / steps
        ranks = per_rank(mine)
        st._graphs.clear()
        return max(ranks), ranks

    use_graph = not args.no_graph
    # ---------------- leg 0: eager, per-kernel CUDA-event timers (roofline of the dominant kernel) --------
    resident = {k: v.to(dev) for k, v in host.items()}
    # the timers are already on during warm-up: with them on, the step stays on one stream, and the
    # caching allocator must have seen exactly the allocation pattern of the timed steps (a first-time
    # cudaMalloc of the 1.1 GB gradient tile buffer inside a timed region read as +0.3 ms of head_grad)
    ops.TIMER.enabled = True
    for _ in range(args.warmup):
        step.micro_step(to_leaves({k: v.detach() for k, v in resident.items()}))
    barrier()
    ops.TIMER.reset()
    n_eager = args.steps if not use_graph else max(20, min(args.steps, 24))   # >= 20 calls per kernel, >= 5 EMA launches
    sampler = ClockSampler(local)
    eager_sampler = ClockSampler(local)
    if not use_graph:
        sampler.start()
    else:
        eager_sampler.start()
    ops.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_eager):
        out = step.micro_step(to_leaves({k: v.detach() for k, v in resident.items()}))
    e1.record()
    barrier()
    eager_clocks = eager_sampler.stop() if use_graph else None
    ops.TIMER.enabled = False
    ops.TIMER.resolve()
    eager_ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / n_eager
    launches = ops.launch_count()
    crops_per_step = sh.student_rows * world

    # ---------------- leg 1: inputs resident in HBM, micro-step replayed as a CUDA graph ----------------
    dev_blobs = [torch.empty(blob_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    bufs = [blob_views(b, feats) for b in dev_blobs]
    if use_graph:
        for p in step.student_head.parameters():
            p.grad = None
        step.micro = 0
        slots = step.static_inputs(feats, slots=2, buffers=bufs)
        for b in range(2):
            dev_blobs[b].copy_(host_blob, non_blocking=True)
        torch.cuda.synchronize()
        step.capture(0)
        step.capture(1)
        for _ in range(args.warmup):
            step.micro_step_graph(0)
        barrier()
        sampler.start()
        e0.record()
        for _ in range(args.steps):
            out = step.micro_step_graph(0)
        e1.record()
        barrier()
        clocks = sampler.stop()
        rank_ms = [x / args.steps for x in per_rank(e0.elapsed_time(e1))]
        ms_per_step = max(rank_ms)
        launches = step.launches_per_graph * args.steps + args.steps // args.accum
    else:
        clocks = sampler.stop()
        ms_per_step = eager_ms_per_step
        rank_ms = [ms_per_step]
    value = crops_per_step / (ms_per_step * 1e-3)
    loss_val = float(out["loss_total"].item())

    # ---------------- leg 1b: the same replay for >= sustained_s seconds (power-limited steady state) -------------
    sustained = None
    if use_graph and args.sustained_s > 0:
        n_sus = max(args.steps, int(math.ceil(args.sustained_s * 1e3 / ms_per_step)))
        sus_sampler = ClockSampler(local)
        barrier()
        sus_sampler.start()
        e0.record()
        for _ in range(n_sus):
            step.micro_step_graph(0)
        e1.record()
        barrier()
        sus_clocks = sus_sampler.stop()
        sus_ms = max_over_ranks(e0.elapsed_time(e1)) / n_sus
        sustained = {"ms_per_step": sus_ms, "value": crops_per_step / (sus_ms * 1e-3), "steps": n_sus,
                     "seconds": sus_ms * n_sus * 1e-3, "clocks": sus_clocks}

    # ---------------- leg 2: end to end from pinned host buffers ----------------
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    read_back = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            dev_blobs[b].copy_(host_blob, non_blocking=True)   # ONE H2D copy per step, straight into the step's inputs
            ready[b].record(copy_stream)

    def e2e_loop(n):
        cur = torch.cuda.current_stream()
        for b in range(2):
            consumed[b].record(cur)
            read_back[b].record(cur)
        prefetch(0)
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                prefetch(i + 1)          # H2D of the next step overlaps this step's kernels
            cur.wait_event(ready[b])
            cur.wait_event(read_back[b])   # the slot's static outputs were read back (two steps ago) before they are rewritten
            if use_graph:
                o = step.micro_step_graph(b)
            else:
                o = step.micro_step(to_leaves({k: v.detach() for k, v in bufs[b].items()}))
            consumed[b].record(cur)
            # D2H of the step's result, every step, on its own stream: a copy queued between two graph launches on the
            # compute stream would put its launch latency on the critical path of every step
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(consumed[b])
                loss_host.copy_(o["loss_total"].reshape(1), non_blocking=True)
                read_back[b].record(d2h_stream)
        cur.synchronize()
        d2h_stream.synchronize()

    e2e_loop(max(args.warmup, 2))
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = crops_per_step / (e2e_ms * 1e-3)

    # ---------------- extra: the other multi-GPU configurations of BASELINE.json at this N ----------------
    # C3 (global batch 256 over 8 GPUs = 32 per GPU, Sinkhorn-Knopp CLS teacher: three all-gathered LSE exchanges per
    # step) and C4 (ViT-L/16, D = 1024, 32 per GPU), per-GPU shapes fixed like the headline (weak scaling).  For C3 the
    # same shapes are replayed WITHOUT the process group as well: the difference is what the collectives cost.
    extra = None
    if use_graph and not args.no_extra and args.config == "C2":
        extra = {}
        step._graphs.clear()
        for name, cfgname, mode in (("C3", "C3", "sinkhorn"), ("C4", "C4", "center")):
            shx = synth.LossHeadShapes(**synth.CONFIGS[cfgname])
            fx = {k: v.to(dev) for k, v in synth.feature_batch(shx, synth.seeded_generator(3, rank), patches_from_tokens=True).items()}
            stx = LossHeadStep(shx, dev, accum=args.accum, process_group=pg, teacher_mode=mode)
            ms_x, ranks_x = replay_leg(stx, fx, args.extra_steps, args.warmup)
            rec = {"workload": f"{cfgname} per-GPU shapes: batch {shx.batch}, D={shx.dim}, K={shx.out_dim}, teacher {mode}, "
                               f"{shx.masked_rows} masked rows, accum {args.accum}, EMA of {stx.n_params / 1e6:.1f} M params",
                   "ms_per_step": ms_x, "value": shx.student_rows * world / (ms_x * 1e-3), "unit": "crops/s",
                   "per_rank_ms": {"min": min(ranks_x), "median": statistics.median(ranks_x), "max": max(ranks_x)},
                   "steps": args.extra_steps}
            if world > 1 and name == "C3":
                del stx
                st0 = LossHeadStep(shx, dev, accum=args.accum, process_group=None, teacher_mode=mode)
                ms_0, _ = replay_leg(st0, fx, args.extra_steps, args.warmup)
                rec["ms_per_step_without_collectives"] = ms_0
                rec["exposed_collective_ms"] = ms_x - ms_0
                stx = st0
            extra[name] = rec
            del stx, fx
            torch.cuda.empty_cache()

    def finish_rank():
        """Teardown: every captured graph (they hold NCCL kernels of the communicator) is destroyed and the device
        idle on EVERY rank before the process group goes away; a watchdog turns a hung communicator teardown into a
        clean exit instead of a stuck job."""
        import gc
        step._graphs.clear()
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            sys.stderr.flush()
            _OUT.flush()
            guard = threading.Timer(30.0, lambda: os._exit(0))
            guard.daemon = True
            guard.start()
            dist.destroy_process_group()
            guard.cancel()

    if rank != 0:
        finish_rank()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peaks = _peaks()
    D, K = sh.dim, sh.out_dim
    rows_s, rows_t = sh.student_rows + sh.masked_rows, sh.teacher_rows + sh.masked_rows
    e_pad = (sh.batch * (sh.n_global * sh.views - sh.n_global) + 127) // 128 * 128 + (sh.masked_rows + 127) // 128 * 128
    # per-kernel time = median over the calls of the eager leg (>= 20 calls; one slow call - a first-time
    # allocation, a host hiccup between the two events - must not read as kernel time)
    kt = {k: float(statistics.median(v)) for k, v in ops.TIMER.samples.items()}
    kn = {k: len(v) for k, v in ops.TIMER.samples.items()}
    if os.environ.get("DINOX_BENCH_DEBUG"):
        for k, v in ops.TIMER.samples.items():
            print(f"[timer] {k:26s}", " ".join(f"{x:.3f}" for x in v), file=sys.stderr)
    from dinox_b200 import losshead as _lh
    readback = _lh.pass2_mode() == "readback"
    # ALGORITHMIC FLOPs per launch: the forward logits each kernel is charged with.  Recompute work is
    # not credited: the statistics passes of the student (recomputed in pass 2) count 0.
    alg_flops = {
        # read-back path: pass 2 forms the student logits of every entry once (the teacher's come from head_teacher)
        "head_grad": 2.0 * D * K * (rows_s if readback else rows_s + rows_t),
        "head_teacher": 2.0 * D * K * rows_t,
        "head_stats_student": 0.0, "head_stats_teacher_cls": 0.0, "head_stats_teacher_patch": 0.0,
        "gemm_dW2": 2.0 * D * K * rows_s, "gemm_dH": 2.0 * D * K * rows_s,
    }
    # ALGORITHMIC HBM bytes per launch of the kernels that stream a rows x K object
    alg_bytes = {
        "head_teacher": 2.0 * rows_t * K + 2.0 * K * D,                     # fp16 probabilities out + W2t in
        "head_grad": (2.0 * rows_t * K if readback else 0.0) + 2.0 * e_pad * K + 2.0 * K * D,   # q in, G out, W2s in
        "gemm_dW2": 2.0 * e_pad * K + 8.0 * K * D, "gemm_dH": 2.0 * e_pad * K + 2.0 * K * D,
        "ema_multi": 12.0 * step.n_params,
    }
    gemm_like = [k for k in kt if alg_flops.get(k, 0.0) > 0.0]
    dom = max(gemm_like or kt, key=kt.get)
    ach = alg_flops.get(dom, 0.0) / (kt[dom] * 1e-3) / 1e12
    # Which measured peak applies: the per-kernel leg is a few tens of milliseconds.  If its own clock record shows the
    # SM clock at (or within 2 % of) the maximum, the chip was in the burst regime -> burst cuBLAS peak; a clock that
    # sagged under the power cap -> sustained peak.  Both fractions are printed.
    ck = eager_clocks if eager_clocks is not None else clocks
    at_max = bool(ck and ck.get("sm_mhz") and ck.get("sm_max_mhz") and ck["sm_mhz"] >= 0.98 * ck["sm_max_mhz"])
    peak_tf = peaks["tf_burst"] if at_max else peaks["tf_sustained"]
    peak_choice = ("burst" if at_max else "sustained") + f" (median SM clock {ck.get('sm_mhz') if ck else None} MHz of " \
        f"{ck.get('sm_max_mhz') if ck else None} during the per-kernel leg)"
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this round
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f)
        k = tj["kernels"].get(dom)
        if k and args.config == "C2" and tj.get("pass2_mode") == _lh.pass2_mode():
            traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
            traffic_src = "profiles/r02_traffic.json (" + tj["source"] + ")"
    except Exception:
        pass
    step_alg_tf = sh.flops() / (ms_per_step * 1e-3) / 1e12
    step_peak = peaks["tf_burst"] if (clocks and clocks.get("sm_mhz") and clocks["sm_mhz"] >= 0.98 * clocks["sm_max_mhz"]) else peaks["tf_sustained"]
    kernels_gbs = {k: alg_bytes[k] / (kt[k] * 1e-3) / 1e9 for k in kt if k in alg_bytes}
    # The roofline that binds the dominant kernel: tensor pipe (algorithmic FLOPs / measured cuBLAS peak) or HBM
    # (algorithmic bytes / measured copy bandwidth), whichever fraction is larger; both are printed.
    frac_tensor = ach / peak_tf
    frac_hbm = kernels_gbs.get(dom, 0.0) / peaks["hbm"]
    hbm_bound = frac_hbm > frac_tensor
    roofline = {
        "kernel": dom, "bound": "hbm" if hbm_bound else "tensor",
        "achieved": kernels_gbs[dom] if hbm_bound else ach, "peak": peaks["hbm"] if hbm_bound else peak_tf,
        "unit": "GB/s" if hbm_bound else "TFLOP/s", "frac": frac_hbm if hbm_bound else frac_tensor,
        "frac_tensor": frac_tensor, "frac_hbm": frac_hbm, "achieved_tflops": ach, "achieved_gbs": kernels_gbs.get(dom),
        "algorithmic_flops_per_launch": alg_flops.get(dom), "algorithmic_bytes_per_launch": alg_bytes.get(dom),
        "frac_of_burst": ach / peaks["tf_burst"], "frac_of_sustained": ach / peaks["tf_sustained"],
        "peak_choice": peak_choice, "peaks": {"bf16_burst": peaks["tf_burst"], "bf16_sustained": peaks["tf_sustained"],
                                              "hbm_gbs": peaks["hbm"], "source": peaks["src"]},
        "traffic": traffic, "traffic_source": traffic_src,
        "kernel_ms": kt[dom], "kernels_ms": kt, "kernels_calls": kn,
        "kernel_ms_stat": f"median over the {kn[dom]} calls of the eager leg",
        "kernels_ms_mean": {k: sum(v) / len(v) for k, v in ops.TIMER.samples.items()},
        "kernels_tflops": {k: alg_flops[k] / (kt[k] * 1e-3) / 1e12 for k in kt if alg_flops.get(k, 0.0) > 0.0},
        "kernels_hbm_gbs": kernels_gbs,
        "kernels_hbm_frac": {k: v / peaks["hbm"] for k, v in kernels_gbs.items()},
        "step_algorithmic_tflops": step_alg_tf, "step_frac": step_alg_tf / step_peak,
        "step_frac_of_burst": step_alg_tf / peaks["tf_burst"], "step_frac_of_sustained": step_alg_tf / peaks["tf_sustained"],
        "ema_gbs": kernels_gbs.get("ema_multi"), "hbm_peak_gbs": peaks["hbm"],
        "eager_leg_clocks": eager_clocks,
    }
    if sustained is not None:
        sus_tf = sh.flops() / (sustained["ms_per_step"] * 1e-3) / 1e12
        sustained["step_algorithmic_tflops"] = sus_tf
        sustained["step_frac_of_sustained_peak"] = sus_tf / peaks["tf_sustained"]
    if world == 1 and not args.no_hbm_table:
        roofline["hbm_kernels"] = hbm_kernel_table(dev, sh, peaks["hbm"])

    torch_eager = None
    if world == 1 and not args.no_torch_eager:
        try:
            step._graphs.clear()
            torch.cuda.empty_cache()
            v, ms = torch_eager_gpu_step_time(args.config, dev, 8, 3, args.accum)
            torch_eager = {"value": v, "unit": "crops/s", "ms_per_step": ms,
                           "what": "the reference algorithm (oracle code) as PyTorch eager ops on this GPU under "
                                   "torch.autocast(bf16): cuBLAS + ATen, logits materialised, per-tensor EMA loop; same "
                                   "workload, inputs resident, 8 steps after 3 warm-up",
                           "speedup_of_this_repo": value / v}
        except Exception as exc:   # e.g. out of memory on a shared GPU: the comparator is optional
            torch_eager = {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, desc, _ = cpu_oracle_step_time(args.config, args.cpu_sample_batch, args.cpu_steps, 2, args.accum,
                                                 args.patch_source == "tokens")
        cpu_baseline = {"value": v, "unit": "crops/s", "cores": cores, "kind": "port", "sample": desc}

    line = {
        "metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": workload_text(args, sh, step.n_params),
            "rows": {"student": sh.student_rows, "teacher": sh.teacher_rows, "masked": sh.masked_rows},
            "parallelism": f"dp{world}", "launch": "cuda-graph replay per micro-step" if use_graph else "eager launches",
            "eager_ms_per_step": eager_ms_per_step, "l2": "per-step working set (bf16 W2 x2 = 100 MB, dL/dlogits 1.1 GB) exceeds the 126 MB L2; no explicit flush",
            "algorithmic_gflop_per_step": sh.flops() / 1e9,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "crops/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roofline,
        "sustained": sustained,
        "per_rank_ms": {"min": min(rank_ms), "median": statistics.median(rank_ms), "max": max(rank_ms)},
        "extra": extra,
        "torch_eager_b200": torch_eager,
        "cpu_baseline": cpu_baseline,
        "loss": loss_val,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    finish_rank()


def run_sweep(args):
    """BASELINE.json configs[4] (C5): K in {8192, 65536, 262144} x N in {196, 576, 1024} at batch 64 per GPU, ViT-S/16
    features, full step (multi-crop CE + iBOT + Gram + EMA/accum).  One JSON line with a `sweep` table: ms per
    micro-step (graph replay, max over ranks), crops/s, algorithmic TFLOP/s, and the per-kernel medians of a short
    eager leg with their fraction of the applicable measured peak."""
    import torch.distributed as dist
    from dinox_b200 import ops, synth
    from dinox_b200.step import LossHeadStep
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = True
    peaks = _peaks()
    rows = []
    for K in synth.C5_SWEEP_K:
        for N in synth.C5_SWEEP_N:
            sh = synth.LossHeadShapes(batch=64, dim=384, out_dim=K, n_patches=N)
            st = LossHeadStep(sh, dev, accum=args.accum, process_group=pg)
            f = {k: v.to(dev) for k, v in synth.feature_batch(sh, synth.seeded_generator(5, rank), patches_from_tokens=True).items()}
            ops.TIMER.enabled = True
            for _ in range(2):
                st.micro_step({k: (v.detach().requires_grad_(True) if k.startswith("student") else v) for k, v in f.items()})
            torch.cuda.synchronize()
            ops.TIMER.reset()
            for _ in range(8):
                st.micro_step({k: (v.detach().requires_grad_(True) if k.startswith("student") else v) for k, v in f.items()})
            torch.cuda.synchronize()
            ops.TIMER.enabled = False
            ops.TIMER.resolve()
            kt = {k: float(statistics.median(v)) for k, v in ops.TIMER.samples.items()}
            for p in st.student_head.parameters():
                p.grad = None
            st.micro = 0
            slots = st.static_inputs(f, slots=1)
            for k, v in f.items():
                slots[0][k].detach().copy_(v)
            st.capture(0)
            for _ in range(3):
                st.micro_step_graph(0)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = args.sweep_steps
            e0.record()
            for _ in range(n):
                st.micro_step_graph(0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            D = sh.dim
            rs, rt = sh.student_rows + sh.masked_rows, sh.teacher_rows + sh.masked_rows
            gram_flops = 2.0 * 2 * sh.teacher_rows * (sh.tokens - 1) ** 2 * D      # student + teacher Gram tiles
            fl = {"head_teacher": 2.0 * D * K * rt, "head_stats_student": 2.0 * D * K * rs, "head_grad": 2.0 * D * K * rs,
                  "gemm_dW2": 2.0 * D * K * rs, "gemm_dH": 2.0 * D * K * rs, "gram_diff": gram_flops}
            rows.append({"K": K, "N": N, "tokens": sh.tokens - 1, "masked_rows": sh.masked_rows, "ms_per_step": ms,
                         "crops_per_s": sh.student_rows * world / (ms * 1e-3),
                         "algorithmic_tflops": sh.flops() / (ms * 1e-3) / 1e12,
                         "frac_of_burst_peak": sh.flops() / (ms * 1e-3) / 1e12 / peaks["tf_burst"],
                         "kernels_ms": kt,
                         "kernels_executed_tflops": {k: fl[k] / (kt[k] * 1e-3) / 1e12 for k in kt if k in fl},
                         "kernels_frac_of_burst_peak": {k: fl[k] / (kt[k] * 1e-3) / 1e12 / peaks["tf_burst"] for k in kt if k in fl}})
            st._graphs.clear()
            del st, f, slots
            torch.cuda.empty_cache()
    if rank == 0:
        line = {"metric": METRIC, "unit": "crops/s", "n_gpus": world, "config": {"workload": "C5 sweep: ViT-S/16 loss head, batch 64 per GPU, "
                "K x N grid, 2 global + 8 local crops, iBOT r=0.3, Gram on, accum %d" % args.accum}, "scaling": "weak",
                "dtype": "bf16", "data": "synthetic", "peaks": peaks, "sweep": rows}
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """Libraries (NCCL prints its version banner) must not add lines to stdout: point fd 1 at stderr
    for the run and hand back a file object on the real stdout for the single JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--accum", type=int, default=4)
    ap.add_argument("--cpu-sample-batch", type=int, default=8)
    ap.add_argument("--teacher-mode", default="center", choices=["center", "sinkhorn"],
                    help="teacher normalisation of the CLS term (C3 uses sinkhorn)")
    ap.add_argument("--cpu-steps", type=int, default=40, help="timed oracle steps of the cpu_baseline leg (~10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--patch-source", default="tokens", choices=["tokens", "rows"],
                    help="iBOT rows: masked rows of the token tensors named by an index (gathered in the step), or rows "
                         "materialised by the caller")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--sustained-s", type=float, default=2.0,
                    help="length of the extra graph-replay leg that reports the power-limited steady state (0 = skip)")
    ap.add_argument("--no-hbm-table", action="store_true", help="skip the GB/s table of the HBM-bound reduction kernels")
    ap.add_argument("--sweep", action="store_true", help="run the C5 sweep grid (K x N) instead of the headline workload")
    ap.add_argument("--sweep-steps", type=int, default=12)
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C4 legs of the `extra` block")
    ap.add_argument("--no-torch-eager", action="store_true", help="skip the PyTorch-eager-on-this-GPU comparator leg")
    ap.add_argument("--no-reference-verbatim", action="store_true",
                    help="--impl reference: skip timing the reference's own classes (only possible where /root/reference exists)")
    ap.add_argument("--extra-steps", type=int, default=24, help="timed micro-steps of each `extra` leg")
    args = ap.parse_args()
    global _OUT
    _OUT = _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.sweep:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --sweep: no CUDA device")
        run_sweep(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the dinox_b200 path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()

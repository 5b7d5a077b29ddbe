"""Graph-replay time of the C2 micro-step with parts removed: what the ~0.5 ms outside the six big GEMM launches
is made of.  Each variant is captured as its own CUDA graph and replayed 60 times (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinox_b200 import synth
from dinox_b200.step import LossHeadStep
dev = torch.device("cuda", 0)
sh = synth.LossHeadShapes(**synth.CONFIGS["C2"])


def run(name, with_tokens=True, with_ibot=True, from_tokens=True, grad=True, gram=True):
    step = LossHeadStep(sh, dev, accum=4)
    f = synth.feature_batch(sh, synth.seeded_generator(2, 0), with_tokens=with_tokens, with_ibot=with_ibot,
                            patches_from_tokens=from_tokens and with_tokens and with_ibot)
    f = {k: v.to(dev) for k, v in f.items()}
    if not gram:
        step.gram_weight = 0.0
    slots = step.static_inputs(f, slots=1)
    for k, v in f.items():
        slots[0][k].data.copy_(v)
        if not grad:
            slots[0][k].requires_grad_(False)
    if not grad:
        for p in step.student_head.parameters():
            p.requires_grad_(False)
        step._fwd_bwd = lambda ff: step._losses(ff)[0]
    step.capture(0)
    for _ in range(5):
        step.micro_step_graph(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 60
    e0.record()
    for _ in range(n):
        step.micro_step_graph(0)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:34s} {e0.elapsed_time(e1) / n:7.3f} ms   launches/graph {step.launches_per_graph}", flush=True)


run("full step")
run("iBOT rows materialised", from_tokens=False)
run("no Gram anchoring (rows)", with_tokens=False, from_tokens=False)
run("no iBOT (CLS + Gram)", with_ibot=False)
run("CLS only", with_tokens=False, with_ibot=False)
run("forward only (full)", grad=False)

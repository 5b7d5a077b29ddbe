"""Per-call samples of the eager timing leg (debug: which calls of a timed region are slow)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinox_b200 import ops, synth
from dinox_b200.step import LossHeadStep
dev = torch.device("cuda", 0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
sh = synth.LossHeadShapes(**synth.CONFIGS[cfg])
step = LossHeadStep(sh, dev, accum=4)
feats = {k: v.to(dev) for k, v in synth.feature_batch(sh, synth.seeded_generator(2, 0)).items()}
leaves = lambda: {k: (v.detach().requires_grad_(True) if k.startswith("student") else v) for k, v in feats.items()}
ops.TIMER.enabled = True
for _ in range(4):
    step.micro_step(leaves())
torch.cuda.synchronize()
ops.TIMER.reset()
for _ in range(8):
    step.micro_step(leaves())
torch.cuda.synchronize()
ops.TIMER.resolve()
print("conc", os.environ.get("DINOX_CONCURRENCY"), cfg)
for k, v in ops.TIMER.samples.items():
    print(f"  {k:26s}", " ".join(f"{x:.3f}" for x in v))

"""CUDA-event table of the HBM-bound reduction kernels of the materialised-logit path (bench.py's `roofline.hbm_kernels`
leg on its own: cold L2 per launch, medians over 20 calls) at one config's logit shapes.  Prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dinox_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
peaks = bench._peaks()
sh = synth.LossHeadShapes(**synth.CONFIGS[cfg])
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
knobs = {k: v for k, v in os.environ.items() if k.startswith("DINOX_")}
print(json.dumps({"config": cfg, "env": knobs, "hbm_peak_gbs": peaks["hbm"],
                  "hbm_kernels": bench.hbm_kernel_table(dev, sh, peaks["hbm"], only=only)}))

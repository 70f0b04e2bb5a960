"""Short workload for ncu: B samples, few timesteps, graphs off so every kernel is a separate launch.

    MDT_GRAPH=0 ncu --metrics gpu__time_duration.sum ... python tools/profile_step.py tf32 4096 3
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import moleculediffusiontransformer_b200 as mdt  # noqa: E402
from bench import MODEL_KW, make_cond  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
model = mdt.QMDiffusion(**MODEL_KW).eval()
out = model.sample(make_cond(B), "cuda:0", cond_scale=7.5, timesteps=steps, seed=1, precision=prec)
torch.cuda.synchronize()
print(float(out.abs().mean()))

"""Parity table of the CUDA path against the reference fixtures (tests/golden): rel-L2 and argmax-token agreement per case and mode.

    python tools/parity_report.py [tf32 bf16 fp32]        # on a B200 (gpurun); prints one line per (case, mode)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import moleculediffusiontransformer_b200 as mdt  # noqa: E402
from oracle import unet_oracle as orc  # noqa: E402
from oracle.cases import CASES, make_inputs  # noqa: E402

modes = sys.argv[1:] or ["tf32"]
models = {}
for name, (kind, kw, mseed, dseed, b, n, cs, steps, clamp) in CASES.items():
    key = (kind, tuple(sorted(kw.items())), mseed)
    if key not in models:
        torch.manual_seed(mseed)
        cls = {"inverse": mdt.QMDiffusion, "forward": mdt.QMDiffusionForward, "analog_sparse": mdt.AnalogDiffusionSparse,
               "analog_full": mdt.AnalogDiffusionFull}[kind]
        models[key] = cls(**kw).eval()
    m = models[key]
    seq, noise0, step_noise = make_inputs(name)
    ref = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))["out"])
    for prec in modes:
        got = m.sample(seq, "cuda:0", cond_scale=cs, timesteps=steps, clamp=clamp, noise=noise0, step_noise=step_noise, precision=prec).cpu()
        agree = (orc.tokens_from_logits(got) == orc.tokens_from_logits(ref)).float().mean().item() if kw["pred_dim"] > 1 else float("nan")
        print(f"{name:24s} {prec:5s} B={b} steps={steps:3d} cs={cs:4.1f}  rel_l2={orc.rel_l2(got, ref):.3e}  tokens={agree:.4f}", flush=True)

"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python -m oracle.make_golden            # all cases of oracle/cases.py

Inputs are regenerated from seeds (oracle/cases.py); weights from torch.manual_seed(model seed).
Only outputs (and a few UNet-level taps) are stored, so fixtures stay a few tens of KB each.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl  # noqa: E402
from oracle.cases import CASES, INPAINT_CASES, make_inpaint_inputs, make_inputs  # noqa: E402


def main(names=None):
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    models = {}
    for name in (names or CASES):
        kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
        key = (kind, tuple(sorted(kw.items())), mseed)
        if key not in models:
            models[key] = rl.build_model(kind, seed=mseed, **kw)
        m = models[key]
        seq, noise0, step_noise = make_inputs(name)
        import contextlib, io
        with rl.injected_noise(noise0, step_noise), contextlib.redirect_stdout(io.StringIO()):
            out = m.sample(seq, "cpu", cond_scale=cs, timesteps=steps, clamp=clamp)
        # one raw UNet evaluation (conditional branch) at sigma = 1.0 for kernel-level parity
        with torch.no_grad():
            x = seq.float().unsqueeze(2)
            emb = m.GELUact(m.fc1(x))
            emb = torch.cat((emb, m.p_enc_1d(emb)), 2)
            t = torch.full((b,), 0.37)
            net = m.unet(noise0, t) if kw.get("unet_type") == "base" else m.unet(noise0, t, embedding=emb, embedding_scale=cs)
        pcount = sum(p.numel() for p in m.unet.parameters()) + m.fc1.weight.numel() + m.fc1.bias.numel()
        psum = float(sum(p.double().sum() for p in {id(p): p for p in m.parameters()}.values()))
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", f"{name}.npz"),
            out=out.numpy(), net=net.numpy(), emb=emb.numpy(), param_sum=np.float64(psum),
            param_count=np.int64(pcount), torch_version=np.bytes_(torch.__version__.encode()))
        print(f"{name}: out.sum={out.double().sum():.6f} net.sum={net.double().sum():.6f} params={pcount}")


def main_inpaint():
    for name, (kw, mseed, dseed, b, n, cs, steps, resamples, keep) in INPAINT_CASES.items():
        m = rl.build_model("inverse", seed=mseed, **kw)
        seq, source, mask, draws = make_inpaint_inputs(name)
        with rl.injected_noise(torch.zeros(0), list(draws)) as st:
            out = m.inpaint(seq, "cpu", cond_scale=cs, timesteps=steps, num_resamples=resamples, inpaint=source, in_paint_mask=mask)
        assert st["i"] == draws.shape[0], (st["i"], draws.shape[0])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{name}.npz"), out=out.numpy())
        print(f"{name}: out.sum={out.double().sum():.6f} draws={st['i']}")


AEULER_CASES = ("inv64_cs7p5", "inv64_short_ctx_clamp")


def main_aeuler():
    """The reference's own injection point: model.diffusion.sample(noise, sampler=AEulerSampler(), ...) with the
    conditioning embedding built exactly as QMDiffusion.sample builds it (generative.py:838-850) and every randn_like injected."""
    rl.load()
    import MoleculeDiffusion.diffusion as rd

    for name in AEULER_CASES:
        kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
        m = rl.build_model(kind, seed=mseed, **kw)
        seq, noise0, step_noise = make_inputs(name)
        with torch.no_grad():
            x = seq.float().unsqueeze(2)
            emb = m.GELUact(m.fc1(x))
            emb = torch.cat((emb, m.p_enc_1d(emb)), 2)
        with rl.injected_noise(noise0, step_noise) as st:
            out = m.diffusion.sample(noise0, embedding=emb, embedding_scale=cs, num_steps=steps, sampler=rd.AEulerSampler(),
                                     sigma_schedule=rd.KarrasSchedule(sigma_min=0.001, sigma_max=9.0, rho=3.0), clamp=clamp)
        assert st["i"] == steps - 1, (st["i"], steps)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"aeuler_{name}.npz"), out=out.numpy())
        print(f"aeuler_{name}: out.sum={out.double().sum():.6f} draws={st['i']}")


# name of the ADPM2 case whose model / inputs are reused -> KarrasSampler kwargs (diffusion.py:404-415)
KARRAS_CASES = {
    "inv64_short_ctx_clamp": dict(s_churn=40.0),                                       # gamma capped at sqrt(2) - 1 on every step
    "inv64_cs7p5": dict(s_churn=10.0, s_tmin=0.05, s_tmax=5.0, s_noise=1.0),           # gamma = 10 / 64 inside the window, 0 outside
}


def main_karras():
    """model.diffusion.sample(noise, sampler=KarrasSampler(s_churn > 0, ...), ...) through the reference's injection point, every
    randn_like injected (one draw per step, diffusion.py:425)."""
    rl.load()
    import MoleculeDiffusion.diffusion as rd

    for name, skw in KARRAS_CASES.items():
        kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
        m = rl.build_model(kind, seed=mseed, **kw)
        seq, noise0, step_noise = make_inputs(name)
        with torch.no_grad():
            x = seq.float().unsqueeze(2)
            emb = m.GELUact(m.fc1(x))
            emb = torch.cat((emb, m.p_enc_1d(emb)), 2)
        with rl.injected_noise(noise0, step_noise) as st:
            out = m.diffusion.sample(noise0, embedding=emb, embedding_scale=cs, num_steps=steps, sampler=rd.KarrasSampler(**skw),
                                     sigma_schedule=rd.KarrasSchedule(sigma_min=0.001, sigma_max=9.0, rho=3.0), clamp=clamp)
        assert st["i"] == steps - 1, (st["i"], steps)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"karras_{name}.npz"), out=out.numpy())
        print(f"karras_{name}: out.sum={out.double().sum():.6f} draws={st['i']}")


if __name__ == "__main__":
    if sys.argv[1:] == ["karras"]:
        main_karras()
    elif sys.argv[1:] == ["inpaint"]:
        main_inpaint()
    elif sys.argv[1:] == ["aeuler"]:
        main_aeuler()
    else:
        main(sys.argv[1:] or None)

"""First-contact GPU diagnostics: per-kernel and per-stage errors against the CPU oracle.

    python tests/gpu_debug.py [fp32|tf32|bf16] [case]        (test infrastructure: it compares against oracle/)
Prints one line per stage tap so a single gpurun call localises a wrong kernel.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import moleculediffusiontransformer_b200 as mdt  # noqa: E402
from moleculediffusiontransformer_b200 import _capi  # noqa: E402
from moleculediffusiontransformer_b200.plan import SamplerPlan  # noqa: E402
from oracle import unet_oracle as orc  # noqa: E402
from oracle.cases import CASES, make_inputs  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def linear_checks(precs):
    lib = _capi.load()
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in [(256, 128, 128), (1000, 512, 256), (4096, 1536, 128), (130, 64, 576), (96, 16, 48), (77, 1, 3)]:
        a = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / K ** 0.5
        bias = torch.randn(N, generator=g)
        res = torch.randn(M, N, generator=g)
        want = torch.nn.functional.gelu(a.double() @ w.double().T + bias.double()) + res.double()
        for prec in precs:
            ad, wd, bd, rd = a.to(dev), w.to(dev), bias.to(dev), res.to(dev)
            out = torch.full((M, N), float("nan"), device=dev)
            rc = lib.mdt_op_linear(ad.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), out.data_ptr(), M, N, K, 1,
                                   _capi.PRECISIONS[prec], None)
            torch.cuda.synchronize()
            if rc != 0:
                print(f"linear {M}x{N}x{K} {prec}: rc={rc} {lib.mdt_last_error().decode()}")
                continue
            print(f"linear {M}x{N}x{K} {prec}: rel={rel(out.cpu().numpy(), want.numpy()):.3e}", flush=True)


def unet_checks(prec, case):
    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[case]
    torch.manual_seed(mseed)
    cls = mdt.QMDiffusion if kind == "inverse" else mdt.QMDiffusionForward
    model = cls(**kw).eval()
    seq, noise0, step_noise = make_inputs(case)
    sd = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith("diffusion.")}
    cfg = model.unet.cfg.to_dict()
    with torch.no_grad():
        emb = orc.encode_conditioning(sd, seq)
        taps_ref = {}
        want_c = orc.unet_forward(sd, cfg, noise0, torch.full((b,), 0.37), emb, taps=taps_ref)
        want = orc.unet_cfg_forward(sd, cfg, noise0, torch.full((b,), 0.37), emb, cs)
    plan = SamplerPlan(model, "cuda:0", precision=prec, max_batch=max(b, 8))
    print(f"plan bytes={plan.device_bytes/1e6:.1f} MB")
    taps = {k: None for k in taps_ref if k != "mapping"}
    got = plan.unet_forward(noise0, 0.37, seq, cond_scale=cs, taps=taps)
    for name, ref in taps_ref.items():
        if name == "mapping" or taps.get(name) is None:
            continue
        r = ref.numpy()  # (b, C, L) -> token-major
        tm = np.transpose(r, (0, 2, 1)).reshape(-1)
        g_ = taps[name][: tm.size]  # conditional half
        print(f"  tap {name:18s} shape={tuple(r.shape)} rel={rel(g_, tm):.3e}", flush=True)
    print(f"unet[{case},{prec}] cond_scale={cs}: rel={rel(got.cpu().numpy(), want.numpy()):.3e}", flush=True)
    return model, plan


def sample_checks(model, prec, case):
    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[case]
    seq, noise0, step_noise = make_inputs(case)
    ref = np.load(os.path.join(ROOT, "tests", "golden", f"{case}.npz"))["out"]
    for graph in ("0", "1"):
        os.environ["MDT_GRAPH"] = graph
        model._plans = {}
        t0 = time.time()
        out = model.sample(seq, "cuda:0", cond_scale=cs, timesteps=steps, clamp=clamp, noise=noise0, step_noise=step_noise,
                           precision=prec)
        torch.cuda.synchronize()
        dt = time.time() - t0
        o = out.cpu()
        agree = (orc.tokens_from_logits(o) == orc.tokens_from_logits(torch.from_numpy(ref))).float().mean().item()
        print(f"sample[{case},{prec},graph={graph}]: rel={rel(o.numpy(), ref):.3e} tokens={agree:.4f} wall={dt:.2f}s", flush=True)


if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    case = sys.argv[2] if len(sys.argv) > 2 else "inv64_cs7p5"
    print(torch.cuda.get_device_name(0))
    linear_checks((sys.argv[3] if len(sys.argv) > 3 else "fp32,tf32,bf16").split(","))
    if case == "none":
        sys.exit(0)
    model, plan = unet_checks(prec, case)
    del plan
    sample_checks(model, prec, case)

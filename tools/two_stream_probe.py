"""Probe: does running two half-batches on two CUDA streams (two plans) beat one full batch on one stream?

    python tools/two_stream_probe.py [precision] [B]
Each half runs its own CUDA graph replay; kernels of the two streams fill each other's tails (wave quantisation, launch gaps)."""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import moleculediffusiontransformer_b200 as mdt  # noqa: E402
from moleculediffusiontransformer_b200 import ADPM2Sampler, KarrasSchedule  # noqa: E402
from moleculediffusiontransformer_b200.plan import SamplerPlan  # noqa: E402
from bench import MODEL_KW, make_cond  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nway = int(sys.argv[3]) if len(sys.argv) > 3 else 2
T = 64
torch.manual_seed(0)
model = mdt.QMDiffusion(**MODEL_KW).eval()
dev = torch.device("cuda", 0)
cond = make_cond(B).to(dev)
sched, sampler = KarrasSchedule(0.001, 9.0, 3.0), ADPM2Sampler(1.0)
kw = dict(num_steps=T, sigma_schedule=sched, sampler=sampler, clamp=False, cond_scale=7.5, return_tokens=False)


def timed(fn, reps=2):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return B * reps / (time.perf_counter() - t0)


one = SamplerPlan(model, dev, precision=prec, max_batch=B)
r1 = timed(lambda: one.sample(cond, seed=1, sample_offset=0, **kw))
ref = one.sample(cond, seed=1, sample_offset=0, **kw).clone()
one.close()

h = B // nway
plans = [SamplerPlan(model, dev, precision=prec, max_batch=h) for _ in range(nway)]
streams = [torch.cuda.Stream(dev) for _ in range(nway)]
outs = [None] * nway


def half(i):
    with torch.cuda.stream(streams[i]):
        outs[i] = plans[i].sample(cond[i * h:(i + 1) * h], seed=1, sample_offset=i * h, **kw)


def both():
    th = [threading.Thread(target=half, args=(i,)) for i in range(nway)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s in streams:
        s.synchronize()


r2 = timed(both)
both()
got = torch.cat(outs)
print(f"{prec} B={B}: one stream {r1:.1f} samples/s, {nway} streams x {h} rows {r2:.1f} samples/s ({r2 / r1:.3f}x); identical={torch.equal(got, ref)}")

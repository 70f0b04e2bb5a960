#!/usr/bin/env python
"""Virtual-screening sweep (BASELINE.json configs[4]): many conditioning vectors, batch-sharded over the GPUs of one box.

    python tools/sweep.py --rows 1000000                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --rows 8000000

Each rank generates its own contiguous block of synthetic U(-1, 1) conditioning rows on the fly (seeded by the global row
index of the chunk, so any sharding produces the same rows), samples with the in-kernel Philox stream keyed by the global
row index, and keeps uint8 tokens; rank 0 gathers them at the end (the only collective of the path)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import moleculediffusiontransformer_b200 as mdt  # noqa: E402
from moleculediffusiontransformer_b200 import ADPM2Sampler, KarrasSchedule  # noqa: E402
from moleculediffusiontransformer_b200.launcher import gather_rows, shard_bounds  # noqa: E402

# README model (BASELINE.json configs[0..1])
INV64 = dict(max_length=64, pred_dim=16, channels=64, unet_type="cfg", context_embedding_max_length=12,
             pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=65536)
    ap.add_argument("--chunk", type=int, default=65536, help="rows generated / sampled per call on each rank")
    ap.add_argument("--cond-scale", type=float, default=5.0)
    ap.add_argument("--timesteps", type=int, default=64)
    ap.add_argument("--precision", default=None, help="default: the package default (fp16 operands)")
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = mdt.QMDiffusion(**INV64).eval()
    lo, hi = shard_bounds(a.rows, world, rank)
    plan = model._plan_for(dev, a.precision, batch=min(a.chunk, hi - lo))
    sched, sampler = KarrasSchedule(0.001, 9.0, 3.0), ADPM2Sampler(1.0)
    toks = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c0 in range(lo, hi, a.chunk):
        c1 = min(hi, c0 + a.chunk)
        # rows are a pure function of their global index: generate in fixed global blocks of `chunk`, slice what this rank owns
        blocks = []
        for g0 in range((c0 // a.chunk) * a.chunk, c1, a.chunk):
            g = torch.Generator().manual_seed(a.seed * 1_000_003 + g0 // a.chunk)
            blk = torch.rand(a.chunk, 12, generator=g) * 2 - 1
            blocks.append(blk[max(c0 - g0, 0): min(c1 - g0, a.chunk)])
        cond = torch.cat(blocks).to(dev)
        _, tok = plan.sample(cond, num_steps=a.timesteps, sigma_schedule=sched, sampler=sampler, clamp=False,
                             cond_scale=a.cond_scale, seed=a.seed, sample_offset=c0, return_tokens=True)
        toks.append(tok)
    local_tokens = torch.cat(toks) if toks else torch.empty((0, 64), dtype=torch.uint8, device=dev)
    full = gather_rows(local_tokens, a.rows) if world > 1 else local_tokens
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"rows={a.rows} gpus={world} time={dt:.2f}s rate={a.rows / dt:.1f} samples/s tokens={tuple(full.shape)} "
              f"checksum={int(full.to(torch.int64).sum())}")
        if a.out:
            torch.save(full.cpu(), a.out)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

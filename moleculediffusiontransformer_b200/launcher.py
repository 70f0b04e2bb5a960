"""Batch-sharded launcher: independent molecules split across the GPUs of one box.

The path is embarrassingly parallel per sample: no collective on the step path.
Each rank processes a contiguous row block of the conditioning matrix with its own plan; the
in-kernel Philox stream is keyed by the GLOBAL sample index, so any sharding of the same
(seed, conditioning) yields identical tokens.  The only communication is one final gather of
uint8 tokens (64 B per molecule) to rank 0, mirroring the single tensor the reference returns.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row block [lo, hi) of `total` rows for `rank` (first ranks take the remainder)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, total: int, group=None, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather ragged row blocks (shard_bounds order) to `dst`; returns the full tensor there, None elsewhere."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(total, ws, r)[1] - shard_bounds(total, ws, r)[0] for r in range(ws)]
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(ws)] if rank == dst else None
    dist.gather(buf, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def sharded_sample(run_shard: Callable[[torch.Tensor, int], torch.Tensor], sequences: torch.Tensor, group=None,
                   dst: int = 0) -> Optional[torch.Tensor]:
    """Run `run_shard(rows, global_row_offset) -> per-row result` on this rank's block and gather to `dst`.

    `sequences` is the full conditioning matrix (every rank may hold it, or just generate its block:
    only rows [lo, hi) are touched)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        ws, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        ws, rank = 1, 0
    lo, hi = shard_bounds(sequences.shape[0], ws, rank)
    local = run_shard(sequences[lo:hi], lo)
    return gather_rows(local, sequences.shape[0], group=group, dst=dst)


def model_runner(model, device, cond_scale: float, timesteps: int, seed: int, precision: Optional[str] = None):
    """run_shard closure over a QMDiffusion / QMDiffusionForward: returns uint8 tokens [rows, L]."""

    def run(rows: torch.Tensor, offset: int) -> torch.Tensor:
        from .diffusion import ADPM2Sampler, KarrasSchedule

        if rows.shape[0] == 0:   # a rank shard_bounds left without rows still takes part in the gather
            return torch.empty((0, model.max_length), dtype=torch.uint8, device=device)
        plan = model._plan_for(torch.device(device), precision, batch=rows.shape[0], timesteps=timesteps)

        _, tokens = plan.sample(rows, num_steps=timesteps, sigma_schedule=KarrasSchedule(0.001, 9.0, 3.0),
                                sampler=ADPM2Sampler(1.0), clamp=False, cond_scale=cond_scale, seed=seed,
                                sample_offset=offset, return_tokens=True)
        return tokens

    return run

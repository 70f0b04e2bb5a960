#!/usr/bin/env python
"""Headline benchmark: generated samples/s of the reverse-diffusion sampling path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision tf32|bf16|fp32] [--batch B]
    python bench.py --impl reference ...          # the reference algorithm on the host CPU cores

One "step" = one full sample() pass (64 timesteps = 63 ADPM2 iterations = 126 denoiser calls, cond_scale 7.5 so
two UNet branches per call) over one batch of B=4096 synthetic 12-property conditioning rows per GPU.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generated samples/sec (64 steps, CFG)"
UNIT = "samples/s"
F_ALG_PER_SAMPLE = 92.2e9      # algorithmic FLOP per generated sample
F_REF_PER_FWD = 455.3e6        # reference-as-executed FLOP per sample per UNet forward
TIMESTEPS, COND_SCALE, N_CTX = 64, 7.5, 12
MODEL_KW = dict(max_length=64, pred_dim=16, channels=64, unet_type="cfg", context_embedding_max_length=12,
                pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)


def workload(batch):
    return {"workload": f"QMDiffusion inverse README arch (ch64, L64, P16, ctx12), batch={batch}/GPU, "
                        f"cond_scale={COND_SCALE}, timesteps={TIMESTEPS}, random init seed 0, U(-1,1) conditioning seed 1",
            "l2_policy": "per-step working set (activation workspace, K/V cache) is several GB >> 126 MB L2; no flush needed"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_cond(batch, offset=0):
    import torch
    g = torch.Generator().manual_seed(1 + offset)
    return torch.rand(batch, N_CTX, generator=g) * 2 - 1


def oracle_bundle():
    import torch
    import moleculediffusiontransformer_b200 as mdt
    torch.manual_seed(0)
    model = mdt.QMDiffusion(**MODEL_KW).eval()
    sd = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith("diffusion.")}
    return model, sd, model.unet.cfg.to_dict()


def cpu_reference_rate(sd, cfg, batch=128, tprime=4, repeats=1):
    """The reference algorithm (oracle port, fp32, all host threads) on a bounded sample, scaled to 64 timesteps.

    Work is linear in ADPM2 iterations (each = 2 denoiser calls x 2 branches), so the full-run rate is
    batch / (t_bounded * (TIMESTEPS - 1) / (tprime - 1))."""
    import torch
    from oracle import unet_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(11)
    seq = torch.rand(batch, N_CTX, generator=g) * 2 - 1
    n0 = torch.randn(batch, 16, 64, generator=g)
    sn = torch.randn(tprime - 1, batch, 16, 64, generator=g)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.sample(sd, cfg, seq, n0, sn, COND_SCALE, tprime, False)
        best = min(best, time.perf_counter() - t0)
    full = best * (TIMESTEPS - 1) / (tprime - 1)
    return batch / full, cores, f"oracle port (plain PyTorch fp32 CPU), batch={batch}, timesteps={tprime} scaled x{(TIMESTEPS - 1) / (tprime - 1):.1f} to 64"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, sd, cfg = oracle_bundle()
    for _ in range(args.warmup):
        cpu_reference_rate(sd, cfg, batch=32, tprime=2)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, cores, sample = cpu_reference_rate(sd, cfg, batch=args.ref_batch, tprime=args.ref_timesteps)
        rates.append(r)
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    v = statistics.mean(rates)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload(args.batch),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_ref_as_executed": v * F_REF_PER_FWD * 252 / 1e9,
    }))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def cublas_peak_tflops(dtype_name):
    """Same method as MEASURED_PEAKS.json (torch.matmul 8192^3, best of 10) for modes it does not list."""
    import torch
    n = 8192
    if dtype_name == "tf32":
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    else:
        torch.backends.cuda.matmul.allow_tf32 = False
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    best = float("inf")
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); (a @ b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = False
    return 2 * n ** 3 / (best * 1e-3) / 1e12


def run_ours(args):
    import torch
    import torch.distributed as dist
    import moleculediffusiontransformer_b200 as mdt
    from moleculediffusiontransformer_b200 import ADPM2Sampler, KarrasSchedule
    from moleculediffusiontransformer_b200.launcher import gather_rows

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = mdt.QMDiffusion(**MODEL_KW).eval()
    B = args.batch
    plan = model._plan_for(dev, args.precision, batch=B)
    sched, sampler = KarrasSchedule(0.001, 9.0, 3.0), ADPM2Sampler(1.0)
    cond_host = make_cond(B, offset=rank).pin_memory()
    cond_dev = cond_host.to(dev)

    def step_resident(i):
        out, tok = plan.sample(cond_dev, num_steps=TIMESTEPS, sigma_schedule=sched, sampler=sampler, clamp=False,
                               cond_scale=COND_SCALE, seed=1234 + i, sample_offset=rank * B, return_tokens=True)
        if world > 1:
            gather_rows(tok, B * world)   # the single collective of the path: final token gather to rank 0
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(100 + i)
    e1.record()
    barrier()
    launches = plan.launch_count - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API: host (pinned) conditioning in, host result out, every step
    def step_e2e():
        out = model.sample(cond_host, dev, cond_scale=COND_SCALE, timesteps=TIMESTEPS, clamp=False, precision=args.precision)
        return out.cpu()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    e2e_value = world * B * args.e2e_steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    if args.precision == "bf16":
        peak, peak_note = peaks["bf16_tflops_sustained"], f"bf16 sustained, {peak_src}"
    elif args.precision == "tf32":
        peak, peak_note = cublas_peak_tflops("tf32"), "cuBLAS TF32 8192^3 best-of-10 measured in this run (MEASURED_PEAKS.json lists bf16 only)"
    else:
        peak, peak_note = cublas_peak_tflops("fp32"), "cuBLAS fp32 (CUDA-core) 8192^3 best-of-10 measured in this run"
    achieved = (value / world) * F_ALG_PER_SAMPLE / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32 operands / f32 accumulate", "bf16": "bf16 operands / f32 accumulate", "fp32": "f32"}[args.precision],
        "data": "synthetic", "config": dict(workload(B), precision=args.precision, parallelism=f"batch-sharded x{world}, no step-path collective"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * N_CTX * 4 + B * 16 * 64 * 4, "d2h_bytes_per_step": B * 16 * 64 * 4,
                "steps": args.e2e_steps},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_note,
                     "note": "algorithmic FLOPs (92.2 GFLOP/sample, BASELINE.md sec. 2) x samples of the timed region / CUDA-event time of the region, per GPU; "
                             "the tcgen05 GEMM kernel is the dominant kernel (share in profiles/)"},
    }
    if args.also and args.also != args.precision:
        # secondary arithmetic mode, same workload, device-resident timing only (reported beside the headline)
        plan2 = model._plan_for(dev, args.also, batch=B)
        def step2(i):
            return plan2.sample(cond_dev, num_steps=TIMESTEPS, sigma_schedule=sched, sampler=sampler, clamp=False,
                                cond_scale=COND_SCALE, seed=1234 + i, sample_offset=rank * B, return_tokens=True)
        for i in range(2):
            step2(i)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(2):
            step2(50 + i)
        f1.record(); torch.cuda.synchronize()
        line["also"] = {"precision": args.also, "value": B * 2 / (f0.elapsed_time(f1) * 1e-3), "unit": UNIT,
                        "note": "same workload in the looser-bound mode (per GPU); not the headline"}
    if world == 1 and not args.no_cpu:
        _, sd, cfg = oracle_bundle()
        v, cores, sample = cpu_reference_rate(sd, cfg, batch=args.ref_batch, tprime=args.ref_timesteps)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MDT_PRECISION", "tf32"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-batch", type=int, default=128)
    ap.add_argument("--ref-timesteps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also", default="bf16", help="secondary precision mode reported under 'also' ('' to skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the reverse-diffusion sampling path: generated samples/s for the BASELINE.json configs.

    python bench.py [--config cfg1|cfg2|cfg3|cfg4|cfg5] [--gpus N] [--steps K] [--warmup W] [--precision tf32|bf16|fp32]
    python bench.py --impl reference ...          # the reference algorithm on the host CPU cores

Default = cfg2, the configuration BASELINE.json's metric is quoted on: one "step" = one full sample() pass (64 timesteps = 63
ADPM2 iterations = 126 denoiser calls, cond_scale 7.5 so two UNet branches per call) over one batch of B = 4096 synthetic
12-property conditioning rows per GPU.  The other configs are BASELINE.json configs[0, 2, 3, 4] measured the same way (their own
batch, model, timesteps and algorithmic FLOP count from BASELINE.md section 2).  Prints ONE JSON line on rank 0.
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

INV64 = dict(max_length=64, pred_dim=16, channels=64, unet_type="cfg", context_embedding_max_length=12,
             pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
FWD64 = dict(INV64, pred_dim=1, context_embedding_max_length=64)
WIDE = dict(INV64, max_length=128, pred_dim=32, channels=128)
MODEL_KW = INV64   # kept for tools/ that import it

# name -> model kind, ctor kwargs, per-GPU batch, context rows, cond_scale, timesteps, algorithmic GFLOP per generated sample
# (BASELINE.md section 2), reference-as-executed MFLOP per sample per UNet forward, UNet forwards per sample(), seed of the inputs
CONFIGS = {
    "cfg1": dict(kind="inverse", kw=INV64, batch=4, n_ctx=12, cond_scale=1.0, timesteps=64, f_alg=49.0e9, f_ref=455.3e6, fwds=126, seed=0,
                 desc="README inverse arch (ch64, L64, P16, ctx12), B=4, cond_scale=1, latency-bound"),
    "cfg2": dict(kind="inverse", kw=INV64, batch=4096, n_ctx=12, cond_scale=7.5, timesteps=64, f_alg=92.2e9, f_ref=455.3e6, fwds=252, seed=1,
                 desc="QMDiffusion inverse README arch (ch64, L64, P16, ctx12)"),
    "cfg3": dict(kind="forward", kw=FWD64, batch=16384, n_ctx=64, cond_scale=1.0, timesteps=64, f_alg=6.7e9, f_ref=223.1e6, fwds=126, seed=2,
                 desc="QMDiffusionForward property predictor (ch64, L64, P1, patch 4, ctx64 synthetic tokenised SMILES / 21)"),
    "cfg4": dict(kind="inverse", kw=WIDE, batch=8192, n_ctx=12, cond_scale=7.5, timesteps=128, f_alg=1025.0e9, f_ref=2228.2e6, fwds=508, seed=3,
                 desc="QMDiffusion inverse widened (ch128, L128, P32, ctx12)"),
    "cfg5": dict(kind="inverse", kw=INV64, batch=65536, n_ctx=12, cond_scale=5.0, timesteps=64, f_alg=92.2e9, f_ref=455.3e6, fwds=252, seed=4,
                 desc="virtual-screening sweep, README arch: bounded sample of the 8M-row sweep (65536 rows per GPU, walked in 4096-row "
                      "chunks exactly as the full sweep is; per-row work and rate do not depend on the row count), uint8 token gather"),
}


def workload(name, batch):
    c = CONFIGS[name]
    return {"workload": f"{name}: {c['desc']}, batch={batch}/GPU, cond_scale={c['cond_scale']}, timesteps={c['timesteps']}, "
                        f"random init seed 0, synthetic conditioning seed {c['seed']}",
            "l2_policy": "per-step working set (activation workspace, K/V cache) is several GB >> 126 MB L2; no flush needed"
                         if batch >= 1024 else "B=4 working set fits L2: latency-bound config, reported as time per call"}


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


def make_cond(batch, offset=0, name="cfg2"):
    """Synthetic conditioning of SURVEY 8(d): U(-1, 1) properties (inverse) or tokenised SMILES / 21, ragged, zero padded (forward)."""
    import torch
    c = CONFIGS[name]
    g = torch.Generator().manual_seed(c["seed"] + offset if name != "cfg2" else 1 + offset)
    n = c["n_ctx"]
    if c["kind"] == "forward":
        lens = torch.randint(1, 30, (batch,), generator=g)
        toks = torch.randint(1, 22, (batch, n), generator=g).float()
        return toks * (torch.arange(n)[None, :] < lens[:, None]) / 21.0
    return torch.rand(batch, n, generator=g) * 2 - 1


def build_model(name):
    import torch
    import moleculediffusiontransformer_b200 as mdt
    c = CONFIGS[name]
    torch.manual_seed(0)
    cls = mdt.QMDiffusion if c["kind"] == "inverse" else mdt.QMDiffusionForward
    return cls(**c["kw"]).eval()


def oracle_bundle(name="cfg2"):
    model = build_model(name)
    sd = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith("diffusion.")}
    return model, sd, model.unet.cfg.to_dict()


def _reference_model(name):
    """The unmodified reference model when its tree is present (this container), else None (the GPU box)."""
    try:
        from oracle import reference_loader as rl
        if not rl.available():
            return None
        c = CONFIGS[name]
        return rl.build_model(c["kind"], seed=0, **c["kw"])
    except Exception:
        return None


def cpu_reference_rate(name, sd, cfg, batch=256, tprime=5, repeats=3, ref_model=None):
    """the reference algorithm, fp32, all host threads, batch min(256, B), reduced timesteps T' >= 5 scaled by
    (T - 1) / (T' - 1) (work is linear in ADPM2 iterations), median of `repeats` timings.  Runs the unmodified reference when its
    tree is importable (kind 'reference'), else the oracle port of the same algorithm (kind 'port')."""
    import torch
    from oracle import unet_oracle as orc
    c = CONFIGS[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = min(batch, c["batch"])
    g = torch.Generator().manual_seed(11)
    seq = make_cond(batch, offset=7, name=name)
    P, L = c["kw"]["pred_dim"], c["kw"]["max_length"]
    n0 = torch.randn(batch, P, L, generator=g)
    sn = torch.randn(tprime - 1, batch, P, L, generator=g)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        if ref_model is not None:
            from oracle import reference_loader as rl
            with rl.injected_noise(n0, list(sn)):
                ref_model.sample(seq, "cpu", cond_scale=c["cond_scale"], timesteps=tprime, clamp=False)
        else:
            orc.sample(sd, cfg, seq, n0, sn, c["cond_scale"], tprime, False)
        times.append(time.perf_counter() - t0)
    T = c["timesteps"]
    full = statistics.median(times) * (T - 1) / (tprime - 1)
    kind = "reference" if ref_model is not None else "port"
    what = "unmodified reference (oracle/reference_loader.py)" if ref_model is not None else "oracle port (plain PyTorch fp32 CPU)"
    return batch / full, cores, kind, (f"{what}, batch={batch}, timesteps={tprime} scaled x{(T - 1) / (tprime - 1):.2f} to {T}, "
                                       f"median of {repeats}")


def linearity_check(name, sd, cfg, ref_model=None):
    """One full-length run at B = 4 against the scaled short run: ratio of per-iteration times."""
    import torch
    from oracle import unet_oracle as orc
    c = CONFIGS[name]
    P, L, T = c["kw"]["pred_dim"], c["kw"]["max_length"], c["timesteps"]
    g = torch.Generator().manual_seed(12)
    seq = make_cond(4, offset=9, name=name)

    def run(steps):
        n0 = torch.randn(4, P, L, generator=g); sn = torch.randn(steps - 1, 4, P, L, generator=g)
        t0 = time.perf_counter()
        if ref_model is not None:
            from oracle import reference_loader as rl
            with rl.injected_noise(n0, list(sn)):
                ref_model.sample(seq, "cpu", cond_scale=c["cond_scale"], timesteps=steps, clamp=False)
        else:
            orc.sample(sd, cfg, seq, n0, sn, c["cond_scale"], steps, False)
        return (time.perf_counter() - t0) / (steps - 1)

    run(3)
    short, full = run(5), run(T)
    return {"per_iteration_s_T5": short, f"per_iteration_s_T{T}": full, "ratio": full / short}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.config
    c = CONFIGS[name]
    _, sd, cfg = oracle_bundle(name)
    ref = _reference_model(name)
    for _ in range(args.warmup):
        cpu_reference_rate(name, sd, cfg, batch=min(32, c["batch"]), tprime=3, repeats=1, ref_model=ref)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, cores, kind, sample = cpu_reference_rate(name, sd, cfg, batch=args.ref_batch, tprime=args.ref_timesteps, repeats=3, ref_model=ref)
        rates.append(r)
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    v = statistics.median(rates)
    lin = linearity_check(name, sd, cfg, ref) if args.linearity else None
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload(name, args.batch or c["batch"]),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_ref_as_executed": v * c["f_ref"] * c["fwds"] / 1e9, "linearity_B4": lin,
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
    torch.backends.cuda.matmul.allow_tf32 = dtype_name == "tf32"
    a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    best = float("inf")
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); (a @ b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = False
    return 2 * n ** 3 / (best * 1e-3) / 1e12


def measured_traffic(name, precision):
    """DRAM bytes per sample() pass from the committed ncu capture (profiles/r02_traffic.json, made by tools/measure_traffic.sh)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(f"{name}_{precision}")
    return (d["dram_bytes_per_sample_pass"], d) if d else (None, None)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from moleculediffusiontransformer_b200 import ADPM2Sampler, KarrasSchedule
    from moleculediffusiontransformer_b200.launcher import gather_rows

    name = args.config
    c = CONFIGS[name]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(name)
    B = args.batch or c["batch"]
    T, CS = c["timesteps"], c["cond_scale"]
    P, L = c["kw"]["pred_dim"], c["kw"]["max_length"]
    plan = model._plan_for(dev, args.precision, batch=B, timesteps=T)
    sched, sampler = KarrasSchedule(0.001, 9.0, 3.0), ADPM2Sampler(1.0)
    cond_host = make_cond(B, offset=rank, name=name).pin_memory()
    cond_dev = cond_host.to(dev)
    tokens = P > 1

    def step_resident(i):
        res = plan.sample(cond_dev, num_steps=T, sigma_schedule=sched, sampler=sampler, clamp=False,
                          cond_scale=CS, seed=1234 + i, sample_offset=rank * B, return_tokens=tokens)
        if world > 1 and tokens:
            gather_rows(res[1], B * world)   # the single collective of the path: final token gather to rank 0
        return res

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
        out = model.sample(cond_host, dev, cond_scale=CS, timesteps=T, clamp=False, precision=args.precision)
        return out.cpu()

    e2e_value = None
    if args.e2e_steps > 0:
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
    if args.precision in ("bf16", "fp16"):
        # kind::f16 runs bf16 and fp16 operands at the same rate: the measured dense bf16 figure is the denominator for both
        peak, peak_note = peaks["bf16_tflops_sustained"], f"bf16 sustained (kind::f16 rate), {peak_src}"
    elif args.precision == "tf32":
        peak, peak_note = cublas_peak_tflops("tf32"), "cuBLAS TF32 8192^3 best-of-10 measured in this run (MEASURED_PEAKS.json lists bf16 only)"
    else:
        peak, peak_note = cublas_peak_tflops("fp32"), "cuBLAS fp32 (CUDA-core) 8192^3 best-of-10 measured in this run"
    achieved = (value / world) * c["f_alg"] / 1e12
    traffic, traffic_detail = measured_traffic(name, args.precision)
    alg_bytes = B * c["fwds"] // (2 if CS != 1.0 else 1) * P * L * 4 * 2   # x in / out per denoiser call (8 KB / sample / call for inv64)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32 operands / f32 accumulate", "bf16": "bf16 operands / f32 accumulate", "fp32": "f32",
                  "fp16": "fp16 operands (11-bit significand, as tf32) / f32 accumulate"}[args.precision],
        "data": "synthetic", "config": dict(workload(name, B), precision=args.precision,
                                            parallelism=f"batch-sharded x{world}, no step-path collective"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * c["n_ctx"] * 4 + B * P * L * 4, "d2h_bytes_per_step": B * P * L * 4,
                "steps": args.e2e_steps},
        "gpu_launches": int(launches), "clocks": clk,
        "ms_per_denoiser_call": ms / args.steps / (2 * (T - 1)), "launches_per_iteration_graph": int(launches) // max(args.steps * (T - 1), 1),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes": alg_bytes,
                     "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None, "traffic_source": traffic_detail,
                     "peak_source": peak_note,
                     "note": f"algorithmic FLOPs ({c['f_alg'] / 1e9:.1f} GFLOP/sample, BASELINE.md sec. 2) x samples of the timed region / CUDA-event "
                             "time of the region, per GPU; one 'launch' = one sample() pass (kernel shares in profiles/); traffic = summed ncu "
                             "dram__bytes of every kernel of one pass"},
    }
    also = []
    # (N = 1 only: under torchrun the other ranks have left by now, and the per-GPU alternatives do not change with N)
    for other in [a for a in args.also.split(",") if a and a != args.precision and world == 1]:
        # secondary arithmetic modes, same workload, device-resident timing only (reported beside the headline)
        plan2 = model._plan_for(dev, other, batch=B, timesteps=T)
        def step2(i):
            return plan2.sample(cond_dev, num_steps=T, sigma_schedule=sched, sampler=sampler, clamp=False,
                                cond_scale=CS, seed=1234 + i, sample_offset=rank * B, return_tokens=tokens)
        for i in range(2):
            step2(i)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(2):
            step2(50 + i)
        f1.record(); torch.cuda.synchronize()
        also.append({"precision": other, "value": B * 2 / (f0.elapsed_time(f1) * 1e-3), "unit": UNIT,
                     "note": "same workload in another arithmetic mode (per GPU); not the headline"})
        for k in [k for k, v in model._plans.items() if v is plan2]:     # free its workspace before the next mode
            del model._plans[k]
        plan2.close()
        del plan2
    if also:
        line["also"] = also
    if world == 1 and not args.no_cpu:
        _, sd, cfg = oracle_bundle(name)
        v, cores, kind, sample = cpu_reference_rate(name, sd, cfg, batch=args.ref_batch, tprime=args.ref_timesteps, repeats=1,
                                                    ref_model=_reference_model(name))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MDT_PRECISION", "fp16"), choices=["fp32", "tf32", "bf16", "fp16"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-batch", type=int, default=256)
    ap.add_argument("--ref-timesteps", type=int, default=5)
    ap.add_argument("--linearity", action="store_true", help="reference arm: add the full-length B=4 linearity check")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also", default="tf32,bf16,fp16", help="other precision modes reported under 'also' (comma separated; '' to skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

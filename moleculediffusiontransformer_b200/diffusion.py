"""Sampler front-end: schedule / sampler strategy objects and the host-side scalar plan.

Mirrors the L3/L2 interface of the reference (diffusion.py) for the one combination the
QM wrappers select -- ``ADPM2Sampler(rho=1)`` + ``KarrasSchedule`` over ``KDiffusion_mod`` --
but none of it runs a Python step loop: the objects only *describe* the run; the loop is a
CUDA-graph-captured driver inside the C-ABI library (csrc/plan.cu).

The scalar plan reproduces the reference's float32-tensor / Python-double mixing exactly:
  KarrasSchedule.forward        diffusion.py:333-342
  ADPM2Sampler.get_sigmas       diffusion.py:495-500   (sigma_up, sigma_down are doubles via math.sqrt,
                                                         sigma_mid stays a float32 0-dim tensor)
  ADPM2Sampler.step             diffusion.py:502-515
  KDiffusion_mod.get_scale_weights / denoise_fn   diffusion.py:789-814
"""
from __future__ import annotations

from math import sqrt
from typing import List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class KarrasSchedule:
    """sigma_i = (smax^(1/rho) + i/(N-1) (smin^(1/rho) - smax^(1/rho)))^rho, padded with 0."""

    def __init__(self, sigma_min: float, sigma_max: float, rho: float = 7.0):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho

    def __call__(self, num_steps: int, device=None) -> torch.Tensor:
        rho_inv = 1.0 / self.rho
        i = torch.arange(num_steps, dtype=torch.float32)
        s = (self.sigma_max ** rho_inv
             + (i / (num_steps - 1)) * (self.sigma_min ** rho_inv - self.sigma_max ** rho_inv)) ** self.rho
        return F.pad(s, pad=(0, 1), value=0.0)


class ADPM2Sampler:
    """DPM-2 ancestral sampler description (diffusion.py:486-524); ``alias`` check as diffusion.py:571-575."""

    diffusion_aliases = ("k", "vk")

    def __init__(self, rho: float = 1.0):
        self.rho = rho

    def get_sigmas(self, sigma: torch.Tensor, sigma_next: torch.Tensor):
        r = self.rho
        sigma_up = sqrt(sigma_next ** 2 * (sigma ** 2 - sigma_next ** 2) / sigma ** 2)
        sigma_down = sqrt(sigma_next ** 2 - sigma_up ** 2)
        sigma_mid = ((sigma ** (1 / r) + sigma_down ** (1 / r)) / 2) ** r
        return sigma_up, sigma_down, sigma_mid


class AEulerSampler:
    """Ancestral Euler sampler description (diffusion.py:456-483): one denoiser call per step.

    On the device it is the second half of an ADPM2 iteration whose midpoint coincides with the start: the scalar table
    carries ``sigma_mid = sigma`` and ``dt_mid = 0`` and the C side skips the first denoiser call for such rows."""

    diffusion_aliases = ("k", "vk")

    def get_sigmas(self, sigma: torch.Tensor, sigma_next: torch.Tensor):
        sigma_up = sqrt(sigma_next ** 2 * (sigma ** 2 - sigma_next ** 2) / sigma ** 2)
        sigma_down = sqrt(sigma_next ** 2 - sigma_up ** 2)
        return sigma_up, sigma_down


class KarrasSampler:
    """Stochastic second-order sampler description (diffusion.py:399-453, "algorithm 2" of arXiv:2206.00364 as the reference
    wrote it): per step  sigma_hat = sigma (1 + gamma),  x_hat = x + sqrt(sigma_hat^2 - sigma^2) s_noise eps,
    d = (x_hat - D(x_hat, sigma_hat)) / sigma_hat,  x' = x_hat + (sigma_next - sigma_hat) d,
    d' = (x' - D(x', sigma_next)) / sigma_next,  x_next = x_hat + 0.5 (sigma - sigma_hat) (d + d')   (diffusion.py:433; the step of
    the correction is zero when s_churn = 0, so a deterministic run returns sigma_0 * noise -- reproduced as is).

    On the device it is an ADPM2-shaped iteration (two denoiser calls) whose second update combines both slopes and whose noise
    is injected ahead of the first call: mdt_plan_set_sampler_mode(plan, 1, ...)."""

    diffusion_aliases = ("k", "vk")

    def __init__(self, s_tmin: float = 0, s_tmax: float = float("inf"), s_churn: float = 0.0, s_noise: float = 1.0):
        self.s_tmin, self.s_tmax, self.s_churn, self.s_noise = s_tmin, s_tmax, s_churn, s_noise

    def gammas(self, sigmas: torch.Tensor, num_steps: int) -> torch.Tensor:
        return torch.where((sigmas >= self.s_tmin) & (sigmas <= self.s_tmax),
                           min(self.s_churn / num_steps, sqrt(2) - 1), 0.0)          # diffusion.py:441-445


# One row per denoiser call / per ADPM2 iteration.  Layout shared with include/mdt_b200.h
# (struct mdt_iter_scalars): 2 x {c_in, c_noise, c_skip, c_out, inv-free sigma divisor} + update coefficients.
ITER_SCALAR_FIELDS = (
    "sigma", "c_in_a", "c_noise_a", "c_skip_a", "c_out_a",
    "sigma_mid", "c_in_b", "c_noise_b", "c_skip_b", "c_out_b",
    "dt_mid", "dt_down", "sigma_up",
)


def _scale_weights(sigma: torch.Tensor, sigma_data: float):
    """KDiffusion_mod.get_scale_weights on a 1-element batch (diffusion.py:789-796)."""
    sigmas = torch.full(size=(1,), fill_value=sigma)  # to_batch, diffusion.py:100
    c_noise = torch.log(sigmas) * 0.25
    s = sigmas.view(1, 1, 1)
    c_skip = (sigma_data ** 2) / (s ** 2 + sigma_data ** 2)
    c_out = s * sigma_data * (sigma_data ** 2 + s ** 2) ** -0.5
    c_in = (s ** 2 + sigma_data ** 2) ** -0.5
    return float(c_in), float(c_noise), float(c_skip), float(c_out)


def karras_noise_scales(sigmas: torch.Tensor, num_steps: int, sampler: "KarrasSampler") -> List[float]:
    """sqrt(sigma_hat^2 - sigma^2) * s_noise of every step (diffusion.py:425-426): math.sqrt of a float32 0-dim tensor."""
    sigmas = sigmas.detach().to("cpu", torch.float32)
    gammas = sampler.gammas(sigmas, num_steps)
    out = []
    for i in range(num_steps - 1):
        sigma_hat = sigmas[i] + gammas[i] * sigmas[i]
        out.append(float(torch.ones((), dtype=torch.float32) * sqrt(sigma_hat ** 2 - sigmas[i] ** 2)) * float(sampler.s_noise))
    return out


def build_iter_scalars(sigmas: torch.Tensor, num_steps: int, sampler, sigma_data: float) -> np.ndarray:
    """Host plan: float32 table [num_steps-1, 13] of every scalar the fused step kernels need."""
    sigmas = sigmas.detach().to("cpu", torch.float32)
    rows: List[List[float]] = []
    if isinstance(sampler, KarrasSampler):
        # rows in the ADPM2 layout: call A at sigma_hat, call B at sigma_next, dt_mid = sigma_next - sigma_hat,
        # dt_down = 0.5 (sigma - sigma_hat) (the factor of d + d'), sigma_up = noise scale of the NEXT step (0 after the last)
        gammas = sampler.gammas(sigmas, num_steps)
        scales = karras_noise_scales(sigmas, num_steps, sampler)
        for i in range(num_steps - 1):
            sig, sig_next = sigmas[i], sigmas[i + 1]
            sig_hat = sig + gammas[i] * sig
            a = _scale_weights(sig_hat, sigma_data)
            b = _scale_weights(sig_next, sigma_data)
            rows.append([float(sig_hat), *a, float(sig_next), *b, float(sig_next - sig_hat), float(0.5 * (sig - sig_hat)),
                         scales[i + 1] if i + 1 < num_steps - 1 else 0.0])
        return np.asarray(rows, dtype=np.float32).reshape(max(num_steps - 1, 0), len(ITER_SCALAR_FIELDS))
    for i in range(num_steps - 1):
        sig, sig_next = sigmas[i], sigmas[i + 1]
        if isinstance(sampler, AEulerSampler):
            sigma_up, sigma_down = sampler.get_sigmas(sig, sig_next)
            sigma_mid = sig                        # degenerate midpoint: the device runs the (x, sigma) evaluation only
        else:
            sigma_up, sigma_down, sigma_mid = sampler.get_sigmas(sig, sig_next)
        sigma_mid = torch.as_tensor(sigma_mid, dtype=torch.float32)
        a = _scale_weights(sig, sigma_data)
        b = _scale_weights(sigma_mid, sigma_data)
        dt_mid = float(sigma_mid - sig)            # diffusion.py:508  (float32 tensor arithmetic)
        dt_down = float(sigma_down - sig)          # diffusion.py:512  (double - float32 tensor -> float32)
        up = float(torch.ones((), dtype=torch.float32) * sigma_up)  # diffusion.py:514 scalar cast
        rows.append([float(sig), *a, float(sigma_mid), *b, dt_mid, dt_down, up])
    return np.asarray(rows, dtype=np.float32).reshape(max(num_steps - 1, 0), len(ITER_SCALAR_FIELDS))


class _KDiffusionShell(nn.Module):
    """Holds ``net`` under the name the reference gives it (KDiffusion_mod, diffusion.py:775-787)."""

    alias = "k"

    def __init__(self, net: nn.Module, sigma_data: float, dynamic_threshold: float = 0.0):
        super().__init__()
        if dynamic_threshold != 0.0:
            raise NotImplementedError("dynamic thresholding (diffusion.py:78-88) is never enabled by the QM wrappers")
        self.net = net
        self.sigma_data = sigma_data
        self.dynamic_threshold = dynamic_threshold


class XDiffusion_x(nn.Module):
    """Front-end shell with the reference's attribute layout (diffusion.py:706-767).

    ``state_dict`` therefore carries ``net.*`` and ``diffusion.net.*`` aliases of the UNet
    parameters exactly like the reference.  ``sample`` is wired by the owning wrapper.
    """

    def __init__(self, type: str, net: nn.Module, *, sigma_data: float, dynamic_threshold: float = 0.0,
                 sigma_distribution=None):
        super().__init__()
        assert type == "k", f"type='{type}' must be 'k' on the accelerated path"
        self.net = net
        self.diffusion = _KDiffusionShell(net, sigma_data, dynamic_threshold)
        self._runner = None  # set by the wrapper: callable(noise, num_steps, schedule, sampler, clamp, **kw)

    def sample(self, noise, num_steps: int, sigma_schedule, sampler, clamp: bool, **kwargs):
        assert self.diffusion.alias in sampler.diffusion_aliases, \
            f"{sampler.__class__.__name__} incompatible with KDiffusion_mod"
        assert num_steps is not None, "Parameter `num_steps` must be provided"
        if self._runner is None:
            raise RuntimeError("XDiffusion_x.sample needs an owning QMDiffusion/QMDiffusionForward wrapper")
        return self._runner(noise=noise, num_steps=num_steps, sigma_schedule=sigma_schedule,
                            sampler=sampler, clamp=clamp, **kwargs)

    def forward(self, *a, **k):  # training loss: out of scope
        raise NotImplementedError("training forward is outside the accelerated sampling path")

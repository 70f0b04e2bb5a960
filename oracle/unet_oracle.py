"""CPU oracle: a plain-PyTorch fp32 restatement of the reference sampling path.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by moleculediffusiontransformer_b200/.

Parity status: PINNED against the reference itself.  The reference has no tests or golden
vectors for this path, so the pin is (1) tests/test_oracle_vs_reference.py,
which runs this file against the unmodified reference imported from /root/reference on the
same weights and injected noise (bit-level agreement expected on CPU), and (2) the committed
fixtures tests/golden/*.npz produced from the reference by oracle/make_golden.py.

Every function cites the reference lines it restates.  It works directly on a state_dict
(name -> tensor) using the reference's key layout; the only configuration it needs is the
small ``cfg`` mapping also consumed by the CUDA library.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- blocks
def conv_block(sd: SD, p: str, x: Tensor, groups: int, scale_shift=None) -> Tensor:
    """ConvBlock1d.forward (modules.py:114-122): GN -> [FiLM] -> SiLU -> conv3."""
    x = F.group_norm(x, groups, sd[p + "groupnorm.weight"], sd[p + "groupnorm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    x = F.silu(x)
    return F.conv1d(x, sd[p + "project.weight"], sd[p + "project.bias"], padding=1)


def resnet_block(sd: SD, p: str, x: Tensor, mapping: Tensor, groups: int) -> Tensor:
    """ResnetBlock1d.forward (modules.py:193-205) with MappingToScaleShift (modules.py:138-142)."""
    h = conv_block(sd, p + "block1.", x, groups)
    ss = F.linear(F.silu(mapping), sd[p + "to_scale_shift.to_scale_shift.1.weight"],
                  sd[p + "to_scale_shift.to_scale_shift.1.bias"])[:, :, None]
    scale, shift = ss.chunk(2, dim=1)
    h = conv_block(sd, p + "block2.", h, groups, (scale, shift))
    if (p + "to_out.weight") in sd:
        x = F.conv1d(x, sd[p + "to_out.weight"], sd[p + "to_out.bias"])
    return h + x


def attention(sd: SD, p: str, x: Tensor, context: Optional[Tensor], heads: int) -> Tensor:
    """Attention.forward + AttentionBase.forward (modules.py:401-410, 350-364)."""
    ctx = x if context is None else context
    xn = F.layer_norm(x, x.shape[-1:], sd[p + "norm.weight"], sd[p + "norm.bias"])
    cn = F.layer_norm(ctx, ctx.shape[-1:], sd[p + "norm_context.weight"], sd[p + "norm_context.bias"])
    q = F.linear(xn, sd[p + "to_q.weight"])
    k, v = F.linear(cn, sd[p + "to_kv.weight"]).chunk(2, dim=-1)
    b, n, hd = q.shape
    d = hd // heads
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    sim = torch.einsum("bhnd,bhmd->bhnm", q, k) * (d ** -0.5)   # scale after QK^T (modules.py:358)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhnm,bhmd->bhnd", attn, v).transpose(1, 2).reshape(b, n, hd)
    return F.linear(out, sd[p + "attention.to_out.weight"], sd[p + "attention.to_out.bias"])


def transformer1d(sd: SD, p: str, x: Tensor, context: Optional[Tensor], heads: int) -> Tensor:
    """Transformer1d.forward (modules.py:519-524); no outer residual."""
    x = F.group_norm(x, 32, sd[p + "to_in.0.weight"], sd[p + "to_in.0.bias"], eps=1e-6)
    x = F.conv1d(x, sd[p + "to_in.1.weight"], sd[p + "to_in.1.bias"]).transpose(1, 2)
    i = 0
    while (p + f"blocks.{i}.attention.to_q.weight") in sd:
        bp = p + f"blocks.{i}."
        x = attention(sd, bp + "attention.", x, None, heads) + x          # modules.py:457
        if (bp + "cross_attention.to_q.weight") in sd:
            x = attention(sd, bp + "cross_attention.", x, context, heads) + x
        h = F.linear(x, sd[bp + "feed_forward.0.weight"], sd[bp + "feed_forward.0.bias"])
        h = F.linear(F.gelu(h), sd[bp + "feed_forward.2.weight"], sd[bp + "feed_forward.2.bias"])
        x = h + x
        i += 1
    x = x.transpose(1, 2)
    return F.conv1d(x, sd[p + "to_out.1.weight"], sd[p + "to_out.1.bias"])


def _count(sd: SD, fmt: str) -> int:
    i = 0
    while any(k.startswith(fmt.format(i)) for k in sd):
        i += 1
    return i


def time_mapping(sd: SD, p: str, time: Tensor) -> Tensor:
    """UNet1d.get_mapping (modules.py:1123-1142) with LearnedPositionalEmbedding (modules.py:554-559)."""
    t = time[:, None]
    freqs = t * sd[p + "to_time.0.0.weights"][None, :] * 2 * math.pi
    four = torch.cat((t, freqs.sin(), freqs.cos()), dim=-1)
    m = F.gelu(F.linear(four, sd[p + "to_time.0.1.weight"], sd[p + "to_time.0.1.bias"]))
    m = F.gelu(F.linear(m, sd[p + "to_mapping.0.weight"], sd[p + "to_mapping.0.bias"]))
    m = F.gelu(F.linear(m, sd[p + "to_mapping.2.weight"], sd[p + "to_mapping.2.bias"]))
    return m


def unet_forward(sd: SD, cfg: dict, x: Tensor, time: Tensor, embedding: Tensor, p: str = "unet.",
                 taps: Optional[dict] = None) -> Tensor:
    """UNet1d.forward (modules.py:1144-1180) for the 'cfg' configuration of the QM wrappers."""
    heads, groups, patch = cfg["attention_heads"], cfg["resnet_groups"], cfg["patch_size"]
    factors = cfg["factors"]

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    mapping = time_mapping(sd, p, time)
    tap("mapping", mapping)
    # Patcher (modules.py:228-231)
    x = resnet_block(sd, p + "to_in.block.", x, mapping, 1)
    if patch > 1:
        b, c, lp = x.shape
        x = x.view(b, c, lp // patch, patch).permute(0, 1, 3, 2).reshape(b, c * patch, lp // patch)
    tap("to_in", x)
    skips_list: List = [x]
    nlev = len(factors)
    for i in range(nlev):
        dp = p + f"downsamples.{i}."
        f = factors[i]
        x = F.conv1d(x, sd[dp + "downsample.weight"], sd[dp + "downsample.bias"], stride=f,
                     padding=f * (cfg["kernel_multiplier_downsample"] // 2))
        tap(f"down{i}.downsample", x)
        skips = []
        if (dp + "pre_transformer_block.to_in.0.weight") in sd:
            x = transformer1d(sd, dp + "pre_transformer_block.", x, None, heads)
            skips.append(x)
            tap(f"down{i}.pre", x)
        for j in range(_count(sd, dp + "blocks.{}.")):
            x = resnet_block(sd, dp + f"blocks.{j}.", x, mapping, groups)
            skips.append(x)
            tap(f"down{i}.res{j}", x)
        if (dp + "transformer.to_in.0.weight") in sd:
            x = transformer1d(sd, dp + "transformer.", x, embedding, heads)
            skips.append(x)
            tap(f"down{i}.tr", x)
        skips_list.append(skips)
    bp = p + "bottleneck."
    x = resnet_block(sd, bp + "pre_block.", x, mapping, groups)
    tap("mid.pre", x)
    if (bp + "transformer.to_in.0.weight") in sd:
        x = transformer1d(sd, bp + "transformer.", x, embedding, heads)
        tap("mid.tr", x)
    x = resnet_block(sd, bp + "post_block.", x, mapping, groups)
    tap("mid.post", x)
    skip_scale = 2 ** -0.5 if cfg.get("use_skip_scale", True) else 1.0
    for u in range(nlev):
        up = p + f"upsamples.{u}."
        skips = skips_list.pop()
        f = factors[nlev - 1 - u]
        for j in range(_count(sd, up + "blocks.{}.")):
            x = torch.cat([x, skips.pop() * skip_scale], dim=1)            # modules.py:829, 844
            x = resnet_block(sd, up + f"blocks.{j}.", x, mapping, groups)
            tap(f"up{u}.res{j}", x)
        if (up + "pre_transformer_block.to_in.0.weight") in sd:
            x = transformer1d(sd, up + "pre_transformer_block.", x, None, heads)
            tap(f"up{u}.pre", x)
        if (up + "transformer.to_in.0.weight") in sd:
            x = transformer1d(sd, up + "transformer.", x, embedding, heads)
            tap(f"up{u}.tr", x)
        x = F.conv_transpose1d(x, sd[up + "upsample.weight"], sd[up + "upsample.bias"], stride=f,
                               padding=f // 2 + f % 2, output_padding=f % 2)
        tap(f"up{u}.upsample", x)
    x = x + skips_list.pop()                                                # modules.py:1176
    # Unpatcher (modules.py:253-257)
    if patch > 1:
        b, cp, l = x.shape
        x = x.view(b, cp // patch, patch, l).permute(0, 1, 3, 2).reshape(b, cp // patch, l * patch)
    x = resnet_block(sd, p + "to_out.block.", x, mapping, 1)
    tap("to_out", x)
    return x


def unet_cfg_forward(sd: SD, cfg: dict, x: Tensor, time: Tensor, embedding: Tensor, embedding_scale: float,
                     p: str = "unet.") -> Tensor:
    """UNetCFG1d.forward (modules.py:1228-1255); a state_dict without ``fixed_embedding`` is the plain UNet1d of
    ``XUNet1d(type='base')`` (modules.py:1144-1180), which the wrappers call without any conditioning (generative.py:862-868)."""
    if (p + "fixed_embedding.embedding.weight") not in sd:
        return unet_forward(sd, cfg, x, time, None, p)
    if embedding_scale != 1.0:
        n = embedding.shape[1]
        fixed = sd[p + "fixed_embedding.embedding.weight"][:n][None].expand(embedding.shape[0], -1, -1)
        out = unet_forward(sd, cfg, x, time, embedding, p)
        out_masked = unet_forward(sd, cfg, x, time, fixed, p)
        return out_masked + (out - out_masked) * embedding_scale
    return unet_forward(sd, cfg, x, time, embedding, p)


# --------------------------------------------------------------------------- wrapper-level pieces
def encode_conditioning(sd: SD, sequences: Tensor, pos_emb_fourier: bool = True, add: bool = False) -> Tensor:
    """generative.py:838-850 + PositionalEncoding1D.forward (transformer.py:3456-3470)."""
    x = sequences.float().unsqueeze(2)
    x = F.gelu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]))
    if pos_emb_fourier:
        inv_freq = sd["p_enc_1d.inv_freq"]
        n, ch = x.shape[1], x.shape[2]
        pos = torch.arange(n, dtype=inv_freq.dtype)
        ang = torch.einsum("i,j->ij", pos, inv_freq)
        emb = torch.cat((ang.sin(), ang.cos()), dim=-1)[None, :, :ch].repeat(x.shape[0], 1, 1)
        x = x + emb if add else torch.cat((x, emb), 2)
    return x


def karras_sigmas(num_steps: int, sigma_min=0.001, sigma_max=9.0, rho=3.0) -> Tensor:
    """KarrasSchedule.forward (diffusion.py:333-342)."""
    rho_inv = 1.0 / rho
    steps = torch.arange(num_steps, dtype=torch.float32)
    s = (sigma_max ** rho_inv + (steps / (num_steps - 1)) * (sigma_min ** rho_inv - sigma_max ** rho_inv)) ** rho
    return F.pad(s, pad=(0, 1), value=0.0)


def denoise(sd: SD, cfg: dict, x_noisy: Tensor, sigma: Tensor, embedding: Tensor, embedding_scale: float,
            sigma_data: float = 0.1) -> Tensor:
    """KDiffusion_mod.denoise_fn (diffusion.py:798-814); x0 clamp always on (dynamic_threshold=0)."""
    b = x_noisy.shape[0]
    sigmas = torch.full(size=(b,), fill_value=sigma)
    c_noise = torch.log(sigmas) * 0.25
    s = sigmas.view(b, 1, 1)
    c_skip = (sigma_data ** 2) / (s ** 2 + sigma_data ** 2)
    c_out = s * sigma_data * (sigma_data ** 2 + s ** 2) ** -0.5
    c_in = (s ** 2 + sigma_data ** 2) ** -0.5
    x_pred = unet_cfg_forward(sd, cfg, c_in * x_noisy, c_noise, embedding, embedding_scale)
    return (c_skip * x_noisy + c_out * x_pred).clamp(-1.0, 1.0)


def adpm2_sample(fn: Callable, noise: Tensor, sigmas: Tensor, num_steps: int, step_noise, rho: float = 1.0) -> Tensor:
    """ADPM2Sampler.forward/step/get_sigmas (diffusion.py:495-524); ``step_noise[i]`` replaces randn_like."""
    x = sigmas[0] * noise
    for i in range(num_steps - 1):
        sigma, sigma_next = sigmas[i], sigmas[i + 1]
        sigma_up = math.sqrt(sigma_next ** 2 * (sigma ** 2 - sigma_next ** 2) / sigma ** 2)
        sigma_down = math.sqrt(sigma_next ** 2 - sigma_up ** 2)
        sigma_mid = ((sigma ** (1 / rho) + sigma_down ** (1 / rho)) / 2) ** rho
        d = (x - fn(x, sigma)) / sigma
        x_mid = x + d * (sigma_mid - sigma)
        d_mid = (x_mid - fn(x_mid, sigma_mid)) / sigma_mid
        x = x + d_mid * (sigma_down - sigma)
        x = x + step_noise[i] * sigma_up
    return x


def aeuler_sample(fn: Callable, noise: Tensor, sigmas: Tensor, num_steps: int, step_noise) -> Tensor:
    """AEulerSampler.forward/step/get_sigmas (diffusion.py:456-483); ``step_noise[i]`` replaces randn_like."""
    x = sigmas[0] * noise
    for i in range(num_steps - 1):
        sigma, sigma_next = sigmas[i], sigmas[i + 1]
        sigma_up = math.sqrt(sigma_next ** 2 * (sigma ** 2 - sigma_next ** 2) / sigma ** 2)
        sigma_down = math.sqrt(sigma_next ** 2 - sigma_up ** 2)
        d = (x - fn(x, sigma)) / sigma
        x_next = x + d * (sigma_down - sigma)
        x = x_next + step_noise[i] * sigma_up
    return x


def karras_sample(fn: Callable, noise: Tensor, sigmas: Tensor, num_steps: int, step_noise, s_tmin: float = 0.0,
                  s_tmax: float = float("inf"), s_churn: float = 0.0, s_noise: float = 1.0) -> Tensor:
    """KarrasSampler.forward/step (diffusion.py:399-453) exactly as written there -- including the second-order correction
    ``x_hat + 0.5 * (sigma - sigma_hat) * (d + d_prime)`` (diffusion.py:433), whose step is zero when s_churn = 0;
    ``step_noise[i]`` replaces the randn_like of step i."""
    x = sigmas[0] * noise
    gammas = torch.where((sigmas >= s_tmin) & (sigmas <= s_tmax), min(s_churn / num_steps, math.sqrt(2) - 1), 0.0)
    for i in range(num_steps - 1):
        sigma, sigma_next, gamma = sigmas[i], sigmas[i + 1], gammas[i]
        sigma_hat = sigma + gamma * sigma
        epsilon = s_noise * step_noise[i]
        x_hat = x + math.sqrt(sigma_hat ** 2 - sigma ** 2) * epsilon
        d = (x_hat - fn(x_hat, sigma_hat)) / sigma_hat
        x_next = x_hat + (sigma_next - sigma_hat) * d
        if sigma_next != 0:
            d_prime = (x_next - fn(x_next, sigma_next)) / sigma_next
            x_next = x_hat + 0.5 * (sigma - sigma_hat) * (d + d_prime)
        x = x_next
    return x


def adpm2_inpaint(fn: Callable, source: Tensor, mask: Tensor, sigmas: Tensor, num_steps: int, num_resamples: int, draws,
                  rho: float = 1.0) -> Tensor:
    """ADPM2Sampler.inpaint (diffusion.py:526-549); ``draws`` replays every randn_like in call order."""
    it = iter(draws)
    x = sigmas[0] * next(it)
    for i in range(num_steps - 1):
        source_noisy = source + sigmas[i] * next(it)
        for r in range(num_resamples):
            x = source_noisy * mask + x * ~mask
            sigma, sigma_next = sigmas[i], sigmas[i + 1]
            sigma_up = math.sqrt(sigma_next ** 2 * (sigma ** 2 - sigma_next ** 2) / sigma ** 2)
            sigma_down = math.sqrt(sigma_next ** 2 - sigma_up ** 2)
            sigma_mid = ((sigma ** (1 / rho) + sigma_down ** (1 / rho)) / 2) ** rho
            d = (x - fn(x, sigma)) / sigma
            x_mid = x + d * (sigma_mid - sigma)
            d_mid = (x_mid - fn(x_mid, sigma_mid)) / sigma_mid
            x = x + d_mid * (sigma_down - sigma)
            x = x + next(it) * sigma_up
            if r < num_resamples - 1:
                s = math.sqrt(sigmas[i] ** 2 - sigmas[i + 1] ** 2)
                x = x + s * next(it)
    return source * mask + x * ~mask


@torch.no_grad()
def inpaint(sd: SD, cfg: dict, sequences: Tensor, source: Tensor, mask: Tensor, draws, cond_scale: float, timesteps: int,
            num_resamples: int = 1) -> Tensor:
    """QMDiffusion.inpaint (generative.py:871-914) -> DiffusionInpainter.forward (diffusion.py:612-625), injected noise."""
    emb = encode_conditioning(sd, sequences)
    sigmas = karras_sigmas(timesteps)
    fn = lambda x, sigma: denoise(sd, cfg, x, sigma, emb, cond_scale)
    return adpm2_inpaint(fn, source, mask, sigmas, timesteps, num_resamples, draws)


@torch.no_grad()
def sample(sd: SD, cfg: dict, sequences: Tensor, noise0: Tensor, step_noise, cond_scale: float, timesteps: int,
           clamp: bool = False, pos_emb_fourier: bool = True, pos_emb_fourier_add: bool = False, sampler: str = "adpm2",
           sampler_kwargs: Optional[dict] = None) -> Tensor:
    """QMDiffusion.sample / QMDiffusionForward.sample (generative.py:834-870, 146-180) with injected noise; ``sampler="aeuler"``
    restates ``model.diffusion.sample(..., sampler=AEulerSampler())`` (diffusion.py:724-741, 456-483)."""
    emb = encode_conditioning(sd, sequences, pos_emb_fourier, pos_emb_fourier_add)
    sigmas = karras_sigmas(timesteps)
    fn = lambda x, sigma: denoise(sd, cfg, x, sigma, emb, cond_scale)
    if sampler == "adpm2":
        x = adpm2_sample(fn, noise0, sigmas, timesteps, step_noise)
    elif sampler == "aeuler":
        x = aeuler_sample(fn, noise0, sigmas, timesteps, step_noise)
    elif sampler == "karras":
        x = karras_sample(fn, noise0, sigmas, timesteps, step_noise, **(sampler_kwargs or {}))
    else:
        raise ValueError(sampler)
    return x.clamp(-1.0, 1.0) if clamp else x


def tokens_from_logits(out: Tensor) -> Tensor:
    """generative.py:1212-1213: permute(0,2,1) -> argmax over the class axis."""
    return out.permute(0, 2, 1).argmax(dim=2)


def rel_l2(a: Tensor, b: Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

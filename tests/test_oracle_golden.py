"""The oracle (oracle/unet_oracle.py) against fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import unet_oracle as orc
from oracle.cases import CASES, INPAINT_CASES, make_inpaint_inputs, make_inputs

FAST = ["inv64_cs1", "inv64_short_ctx_clamp", "fwd64_cs2", "paper_cs2", "base128_inv", "base128_fwd", "analog_sparse", "analog_full"]


def _sd_cfg(model):
    sd = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith("diffusion.")}
    return sd, model.unet.cfg.to_dict()


@pytest.mark.parametrize("name", list(CASES))
def test_seeded_init_matches_reference(name, model_cache):
    kind, kw, mseed, *_ = CASES[name]
    m = model_cache(kind, kw, mseed)
    g = golden(name)
    uniq = {id(p): p for p in m.parameters()}
    assert sum(p.numel() for p in uniq.values()) == int(g["param_count"])
    psum = float(sum(p.detach().double().sum() for p in uniq.values()))
    assert psum == pytest.approx(float(g["param_sum"]), rel=0, abs=1e-9)


@pytest.mark.parametrize("name", list(CASES))
def test_unet_eval_matches_reference(name, model_cache):
    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
    m = model_cache(kind, kw, mseed)
    sd, cfg = _sd_cfg(m)
    seq, noise0, _ = make_inputs(name)
    g = golden(name)
    with torch.no_grad():
        emb = orc.encode_conditioning(sd, seq)
        assert np.array_equal(emb.numpy(), g["emb"])
        net = orc.unet_cfg_forward(sd, cfg, noise0, torch.full((b,), 0.37), emb, cs)
    # identical op sequence on the same CPU kernels: expect (near) bit equality
    assert orc.rel_l2(net, torch.from_numpy(g["net"])) < 1e-6


@pytest.mark.parametrize("name", FAST)
def test_full_sample_matches_reference(name, model_cache):
    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
    m = model_cache(kind, kw, mseed)
    sd, cfg = _sd_cfg(m)
    seq, noise0, step_noise = make_inputs(name)
    out = orc.sample(sd, cfg, seq, noise0, step_noise, cs, steps, clamp)
    ref = torch.from_numpy(golden(name)["out"])
    assert out.shape == ref.shape == (b, kw["pred_dim"], kw["max_length"])
    assert orc.rel_l2(out, ref) < 1e-5
    assert (orc.tokens_from_logits(out) == orc.tokens_from_logits(ref)).float().mean() == 1.0


def test_survey_self_check_values():
    """SURVEY.md 8(c) self-check numbers for the README model (seed 0, generator 1234)."""
    g = golden("inv64_cs1")["out"]
    assert float(g.astype(np.float64).sum()) == pytest.approx(191.704454, abs=2e-4)
    assert np.allclose(g[0, 0, :4], [0.092168, 0.032204, 0.119221, 0.249968], atol=2e-6)
    toks = g.transpose(0, 2, 1).argmax(2)
    assert toks[0, :16].tolist() == [8, 11, 1, 11, 9, 13, 14, 5, 1, 9, 9, 3, 11, 0, 11, 3]
    g2 = golden("inv64_cs7p5")["out"]
    assert float(g2.astype(np.float64).sum()) == pytest.approx(289.317932, abs=4e-4)


@pytest.mark.parametrize("name", list(INPAINT_CASES))
def test_inpaint_matches_reference(name, model_cache):
    """Oracle restatement of ADPM2Sampler.inpaint against fixtures produced by the reference's QMDiffusion.inpaint."""
    kw, mseed, dseed, b, n, cs, steps, resamples, keep = INPAINT_CASES[name]
    m = model_cache("inverse", kw, mseed)
    sd, cfg = _sd_cfg(m)
    seq, source, mask, draws = make_inpaint_inputs(name)
    out = orc.inpaint(sd, cfg, seq, source, mask, list(draws), cs, steps, resamples)
    ref = torch.from_numpy(golden(name)["out"])
    assert orc.rel_l2(out, ref) < 1e-5
    assert torch.equal(out[:, :, :keep], source[:, :, :keep])          # kept region is the draft, bit for bit


@pytest.mark.parametrize("name", ["inv64_cs7p5", "inv64_short_ctx_clamp"])
def test_aeuler_sampler_matches_reference(name, model_cache):
    """Oracle restatement of AEulerSampler (diffusion.py:456-483) against fixtures produced by the reference through its own
    injection point ``model.diffusion.sample(noise, sampler=AEulerSampler(), ...)`` (oracle/make_golden.py aeuler)."""
    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
    m = model_cache(kind, kw, mseed)
    sd, cfg = _sd_cfg(m)
    seq, noise0, step_noise = make_inputs(name)
    out = orc.sample(sd, cfg, seq, noise0, step_noise, cs, steps, clamp, sampler="aeuler")
    ref = torch.from_numpy(golden("aeuler_" + name)["out"])
    assert orc.rel_l2(out, ref) < 1e-5
    assert (orc.tokens_from_logits(out) == orc.tokens_from_logits(ref)).float().mean() == 1.0
    assert orc.rel_l2(out, torch.from_numpy(golden(name)["out"])) > 1e-2        # and it is not the ADPM2 result


@pytest.mark.parametrize("name", ["inv64_short_ctx_clamp", "inv64_cs7p5"])
def test_karras_sampler_matches_reference(name, model_cache):
    """Oracle restatement of KarrasSampler with s_churn > 0 (diffusion.py:399-453) against fixtures produced by the reference
    through ``model.diffusion.sample(noise, sampler=KarrasSampler(...), ...)`` (oracle/make_golden.py karras)."""
    from oracle.make_golden import KARRAS_CASES

    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
    m = model_cache(kind, kw, mseed)
    sd, cfg = _sd_cfg(m)
    seq, noise0, step_noise = make_inputs(name)
    out = orc.sample(sd, cfg, seq, noise0, step_noise, cs, steps, clamp, sampler="karras", sampler_kwargs=KARRAS_CASES[name])
    ref = torch.from_numpy(golden("karras_" + name)["out"])
    assert orc.rel_l2(out, ref) < 1e-5
    assert orc.rel_l2(out, torch.from_numpy(golden(name)["out"])) > 1e-2        # and it is not the ADPM2 result


def test_karras_scalar_rows_follow_the_reference_arithmetic():
    """Host plan rows for KarrasSampler: sigma_hat, the two coefficient sets, dt_mid, 0.5 (sigma - sigma_hat) and the next step's
    noise scale, each computed with the reference's float32-tensor / Python-double mix (diffusion.py:422-434, 441-445)."""
    import math
    import moleculediffusiontransformer_b200 as mdt
    from moleculediffusiontransformer_b200.diffusion import karras_noise_scales

    N = 9
    sig = orc.karras_sigmas(N)
    smp = mdt.KarrasSampler(s_churn=3.0, s_tmin=0.01, s_tmax=4.0, s_noise=1.0)
    rows = mdt.build_iter_scalars(sig, N, smp, 0.1)
    scales = karras_noise_scales(sig, N, smp)
    gam = torch.where((sig >= 0.01) & (sig <= 4.0), min(3.0 / N, math.sqrt(2) - 1), 0.0)
    assert rows.shape == (N - 1, 13) and float(gam[0]) == 0.0 and float(gam[3]) > 0
    for i in range(N - 1):
        s_hat = sig[i] + gam[i] * sig[i]
        assert rows[i, 0] == float(s_hat) and rows[i, 5] == float(sig[i + 1])
        assert rows[i, 10] == float(sig[i + 1] - s_hat) and rows[i, 11] == float(0.5 * (sig[i] - s_hat))
        assert scales[i] == float(torch.ones(()) * math.sqrt(s_hat ** 2 - sig[i] ** 2))
        assert rows[i, 12] == (np.float32(scales[i + 1]) if i + 1 < N - 1 else 0.0)

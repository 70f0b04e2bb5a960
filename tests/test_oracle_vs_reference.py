"""Pin the oracle to the unmodified reference, live, when /root/reference exists (build container only)."""
import pytest
import torch

from oracle import reference_loader as rl
from oracle import unet_oracle as orc
from oracle.cases import CASES, make_inputs

pytestmark = pytest.mark.skipif(not rl.available(), reason="reference tree not present (GPU box)")


@pytest.mark.parametrize("name", ["inv64_short_ctx_clamp", "fwd64_cs2"])
def test_sample_live(name):
    kind, kw, mseed, dseed, b, n, cs, steps, clamp = CASES[name]
    ref = rl.build_model(kind, seed=mseed, **kw)
    seq, noise0, step_noise = make_inputs(name)
    with rl.injected_noise(noise0, step_noise) as st:
        want = ref.sample(seq, "cpu", cond_scale=cs, timesteps=steps, clamp=clamp)
    assert st["i"] == steps - 1 and st["used0"]            # RNG draw counts
    sd = {k: v.detach() for k, v in ref.state_dict().items() if not k.startswith("diffusion.")}
    cfg = dict(attention_heads=8, resnet_groups=8, patch_size=1 if kind == "inverse" else 4, factors=[4, 4],
               kernel_multiplier_downsample=2, use_skip_scale=True)
    got = orc.sample(sd, cfg, seq, noise0, step_noise, cs, steps, clamp)
    assert orc.rel_l2(got, want) < 1e-6


def test_state_dict_layout_and_init_live():
    import moleculediffusiontransformer_b200 as mdt

    kind, kw, mseed, *_ = CASES["fwd64_cs1"]
    ref = rl.build_model(kind, seed=mseed, **kw)
    torch.manual_seed(mseed)
    mine = mdt.QMDiffusionForward(**kw)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)
    # triple aliasing: unet.* == diffusion.net.* == diffusion.diffusion.net.* share storage
    k = "unet.to_mapping.0.weight"
    assert b[k].data_ptr() == b["diffusion.net." + k[5:]].data_ptr() == b["diffusion.diffusion.net." + k[5:]].data_ptr()
    ref.load_state_dict(b, strict=True)
    mine.load_state_dict(a, strict=True)


def test_training_forward_delegate_live():
    """set_training_delegate: the reference object computes the training loss (generative.py:812-833) on this model's weights."""
    import moleculediffusiontransformer_b200 as mdt
    kind, kw, mseed, *_ = CASES["inv64_cs1"]
    torch.manual_seed(3)
    mine = mdt.QMDiffusion(**kw)
    ref = rl.build_model(kind, seed=7, **kw)                 # different weights: the delegate must take ours
    mine.set_training_delegate(ref)
    g = torch.Generator().manual_seed(1)
    seq, out = torch.rand(2, 12, generator=g), torch.rand(2, 16, 64, generator=g) * 2 - 1
    torch.manual_seed(11)
    loss = mine(seq, out)
    ref2 = rl.build_model(kind, seed=9, **kw)
    ref2.load_state_dict(mine.state_dict(), strict=True)
    torch.manual_seed(11)
    want = ref2(seq, out)
    assert loss.dim() == 0 and torch.equal(loss, want)

"""CPU-only checks: C-ABI library loads and exports every declared symbol, host scalar plan, launcher logic."""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from moleculediffusiontransformer_b200 import _capi

    lib = _capi.load()
    header = open(os.path.join(ROOT, "include", "mdt_b200.h")).read()
    declared = set(re.findall(r"\b(mdt_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.mdt_abi_version() == _capi.MDT_ABI_VERSION


def test_no_cpu_fallback():
    import moleculediffusiontransformer_b200 as mdt
    from oracle.cases import INV64

    torch.manual_seed(0)
    m = mdt.QMDiffusion(**INV64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.sample(torch.zeros(2, 12), "cpu", cond_scale=1.0, timesteps=4)
    with pytest.raises(RuntimeError):
        m.unet(torch.zeros(1, 16, 64))          # parameter containers have no eager forward
    if not torch.cuda.is_available():
        from moleculediffusiontransformer_b200 import _capi
        assert _capi.load().mdt_device_count() == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "moleculediffusiontransformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"


def test_iter_scalars_match_reference_arithmetic():
    """build_iter_scalars vs the oracle's restatement of ADPM2Sampler.get_sigmas / get_scale_weights."""
    from moleculediffusiontransformer_b200 import ADPM2Sampler, KarrasSchedule, build_iter_scalars
    from oracle import unet_oracle as orc

    for n in (2, 5, 64, 100):
        sig = KarrasSchedule(0.001, 9.0, 3.0)(n)
        assert torch.equal(sig, orc.karras_sigmas(n))
        tab = build_iter_scalars(sig, n, ADPM2Sampler(1.0), 0.1)
        assert tab.shape == (n - 1, 13)
        for i in range(n - 1):
            s, sn = sig[i], sig[i + 1]
            up = math.sqrt(sn ** 2 * (s ** 2 - sn ** 2) / s ** 2)
            down = math.sqrt(sn ** 2 - up ** 2)
            mid = ((s ** 1.0 + down ** 1.0) / 2) ** 1.0
            assert tab[i, 0] == np.float32(s) and tab[i, 5] == np.float32(mid)
            assert tab[i, 10] == np.float32(mid - s) and tab[i, 11] == np.float32(down - s)
            sigmas = torch.full((1,), s)
            assert tab[i, 2] == np.float32(torch.log(sigmas) * 0.25)
            assert tab[i, 1] == np.float32((sigmas ** 2 + 0.1 ** 2) ** -0.5)


def test_aeuler_rows_are_degenerate_adpm2_rows():
    """AEulerSampler rows: midpoint at the start (what the C side recognises as one denoiser call per step), same sigma_up /
    sigma_down as ADPM2 (diffusion.py:467-469 vs 495-498)."""
    from moleculediffusiontransformer_b200 import ADPM2Sampler, AEulerSampler, KarrasSchedule, build_iter_scalars

    sig = KarrasSchedule(0.001, 9.0, 3.0)(16)
    ae = build_iter_scalars(sig, 16, AEulerSampler(), 0.1)
    ad = build_iter_scalars(sig, 16, ADPM2Sampler(1.0), 0.1)
    assert ae.shape == ad.shape == (15, 13)
    assert np.array_equal(ae[:, 5], ae[:, 0]) and np.all(ae[:, 10] == 0.0)                  # sigma_mid == sigma, dt_mid == 0
    assert np.array_equal(ae[:, 6:10], ae[:, 1:5])                                          # call-B coefficients are sigma's
    assert np.array_equal(ae[:, [0, 1, 2, 3, 4, 11, 12]], ad[:, [0, 1, 2, 3, 4, 11, 12]])   # sigma, call-A coefficients, dt_down, sigma_up
    assert np.all(ad[:, 10] < 0.0)                                                          # an ADPM2 row can never look first order


def test_c_scalar_helpers_match_python():
    from moleculediffusiontransformer_b200 import ADPM2Sampler, KarrasSchedule, _capi, build_iter_scalars

    lib = _capi.load()
    for n in (8, 64, 128):
        sig = KarrasSchedule(0.001, 9.0, 3.0)(n)
        c = np.zeros(n + 1, np.float32)
        _capi.check(lib.mdt_karras_sigmas(n, 0.001, 9.0, 3.0, c.ctypes.data))
        assert np.allclose(c[:-1], sig.numpy()[:-1], rtol=3e-7) and c[-1] == 0
        tab = build_iter_scalars(sig, n, ADPM2Sampler(1.0), 0.1)
        arr = (_capi.MdtIterScalars * (n - 1))()
        s32 = sig.numpy().copy()
        _capi.check(lib.mdt_adpm2_scalars(s32.ctypes.data, n - 1, 1.0, 0.1, arr))
        ct = np.ctypeslib.as_array(ctypes.cast(arr, ctypes.POINTER(ctypes.c_float)), shape=(n - 1, 13))
        assert np.allclose(ct, tab, rtol=5e-7, atol=0)
        from moleculediffusiontransformer_b200 import AEulerSampler
        tab_e = build_iter_scalars(sig, n, AEulerSampler(), 0.1)
        arr_e = (_capi.MdtIterScalars * (n - 1))()
        _capi.check(lib.mdt_aeuler_scalars(s32.ctypes.data, n - 1, 0.1, arr_e))
        ce = np.ctypeslib.as_array(ctypes.cast(arr_e, ctypes.POINTER(ctypes.c_float)), shape=(n - 1, 13))
        assert np.allclose(ce, tab_e, rtol=5e-7, atol=0)
        assert np.array_equal(ce[:, 5], ce[:, 0]) and np.all(ce[:, 10] == 0.0)      # exactly the rows the driver recognises
    assert lib.mdt_karras_sigmas(1, 0.001, 9.0, 3.0, c.ctypes.data) < 0
    assert b"num_steps" in lib.mdt_last_error()


def test_plan_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import moleculediffusiontransformer_b200 as mdt
    from moleculediffusiontransformer_b200.plan import SamplerPlan
    from oracle.cases import INV64

    torch.manual_seed(0)
    m = mdt.QMDiffusion(**INV64)
    with pytest.raises(RuntimeError):
        SamplerPlan(m, "cuda:0")


def test_shard_bounds_cover_and_balance():
    from moleculediffusiontransformer_b200.launcher import shard_bounds

    for total in (0, 1, 7, 4096, 8_000_003):
        for ws in (1, 2, 4, 8):
            spans = [shard_bounds(total, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from moleculediffusiontransformer_b200.launcher import sharded_sample
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
total = 11
seq = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
def run(rows, offset):  # stand-in for the CUDA plan: a pure function of (global row index, row content)
    idx = torch.arange(offset, offset + rows.shape[0])
    return (idx[:, None] * 7 + rows.sum(1, keepdim=True).long() + torch.arange(5)[None]).to(torch.uint8)
full = sharded_sample(run, seq)
if dist.get_rank() == 0:
    want = run(seq, 0)
    assert full is not None and torch.equal(full, want), (full, want)
    print("OK")
else:
    assert full is None
dist.destroy_process_group()
"""


def test_sharded_gather_world2_gloo(tmp_path):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0][0]


def test_tokens_to_forward_conditioning_matches_text_round_trip():
    """Compaction of padding + normalisation == decode to text, re-tokenise, divide, pad (generative.py:682-685, 1069-1078)."""
    from moleculediffusiontransformer_b200.screening import tokens_to_forward_conditioning

    toks = torch.tensor([[3, 0, 5, 5, 0, 0, 1, 2], [0, 0, 0, 0, 0, 0, 0, 0], [7, 6, 5, 4, 3, 2, 1, 9]], dtype=torch.uint8)
    got = tokens_to_forward_conditioning(toks, 6, 21.0)
    want = torch.zeros(3, 6)
    for i, row in enumerate(toks.tolist()):
        ids = [t for t in row if t != 0][:6]
        want[i, : len(ids)] = torch.tensor(ids, dtype=torch.float32) / 21.0
    assert torch.equal(got, want)


QM9_LIKE_VOCAB = {i + 1: ch for i, ch in enumerate("CNOF()=#123456[]+-Hcno")}   # char-level SMILES tokeniser, 22 ids like the paper's


def test_decode_oracle_follows_keras_sequences_to_texts():
    """reverse_tokenize (generative.py:1069-1078): ids without a vocabulary entry (padding 0 among them) vanish, spaces stripped."""
    from oracle.decode_oracle import reverse_tokenize, sequences_to_texts

    assert sequences_to_texts({1: "C", 2: "N"}, [[1, 2, 0, 1], [0, 0], [9, 1]]) == ["C N C", "", "C"]
    x = np.array([[1, 1, 3, 0, 0], [0, 5, 0, 7, 200], [0, 0, 0, 0, 0]]) / 21.0
    assert reverse_tokenize(QM9_LIKE_VOCAB, x, 21.0) == ["CCO", "(=", ""]


def _decode_record():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "decode_notebook.json")))


def test_decode_oracle_against_notebook_record():
    """The reference-held known answers for reverse_tokenize (generative.py:1069-1078): vocabulary, token rows and the strings the
    authors' own run printed (Inverse_Diffusion.ipynb cells 36 / 38 / 65, extracted by oracle/make_decode_fixture.py)."""
    import numpy as np
    from oracle.decode_oracle import reverse_tokenize

    rec = _decode_record()
    index_word = {int(k): v for k, v in rec["index_word"].items()}
    assert {v: k for k, v in index_word.items()} == rec["word_index"] and len(index_word) == 21
    rows = np.asarray(rec["tokenized_rows"])
    assert reverse_tokenize(index_word, rows) == rec["reverse_tokenized"]
    # the call site divides by X_norm_factor first and reverse_tokenize multiplies it back (generative.py:1071, 1207-1229)
    assert reverse_tokenize(index_word, rows / rec["x_norm_factor"], rec["x_norm_factor"]) == rec["reverse_tokenized"]
    # strings the reference decoded from sampled tokens: tokenise with the recorded vocabulary, pad with id 0, decode
    for smi in rec["decoded_smiles"]:
        ids = [rec["word_index"][ch] for ch in smi] + [0] * (32 - len(smi))
        assert reverse_tokenize(index_word, np.asarray([ids])) == [smi]
        scattered = np.zeros(64, dtype=np.int64); scattered[1:2 * len(smi):2] = ids[:len(smi)]   # padding between characters is dropped
        assert reverse_tokenize(index_word, scattered[None]) == [smi]


def test_vocabulary_table_rejects_what_the_device_decoder_cannot_express():
    from moleculediffusiontransformer_b200.screening import is_novel, vocabulary_table

    lut = vocabulary_table(QM9_LIKE_VOCAB)
    assert lut.shape == (256,) and lut[0] == 0 and chr(lut[1]) == "C" and lut[23] == 0
    for bad in ({0: "C"}, {256: "C"}, {3: "Cl"}, {3: " "}, {3: "é"}):
        with pytest.raises(ValueError):
            vocabulary_table(bad)
    assert is_novel(["CCO"], "CCN") and not is_novel(["CCO"], "CCO")


def test_empty_shard_and_plan_cache_housekeeping():
    """A rank that shard_bounds leaves without rows must still return a [0, L] uint8 block for the gather (no CUDA call is made);
    invalidate_plans() / the weights fingerprint exist for writes that bypass PyTorch's version counter."""
    import torch
    import moleculediffusiontransformer_b200 as mdt
    from moleculediffusiontransformer_b200.launcher import model_runner, shard_bounds

    torch.manual_seed(0)
    m = mdt.QMDiffusion(max_length=64, pred_dim=16, channels=64, unet_type="cfg", context_embedding_max_length=12,
                        pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
    assert shard_bounds(1, 2, 1) == (1, 1)
    tok = model_runner(m, "cpu", 5.0, 6, seed=1)(torch.zeros(0, 12), 1)
    assert tok.shape == (0, 64) and tok.dtype == torch.uint8
    f0 = m._weights_fingerprint()
    with torch.no_grad():
        m.fc1.weight.add_(1.0)                      # bumps the version counter
    assert m._weights_fingerprint() != f0
    m.invalidate_plans()
    assert m._plans == {}


def test_training_forward_delegates_to_a_reference_object():
    """forward(sequences, output) is outside the accelerated path: with a delegate attached it runs there on this model's weights."""
    import torch
    import moleculediffusiontransformer_b200 as mdt

    torch.manual_seed(0)
    kw = dict(max_length=64, pred_dim=16, channels=64, unet_type="cfg", context_embedding_max_length=12,
              pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
    m = mdt.QMDiffusion(**kw)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(2, 12), torch.zeros(2, 16, 64))

    class FakeReference(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.loaded = None
        def load_state_dict(self, sd, *a, **k):
            self.loaded = {k2: v.clone() for k2, v in sd.items()}
        def forward(self, sequences, output):
            return output.mean() + self.loaded["fc1.bias"].sum() * 0

    ref = FakeReference()
    m.set_training_delegate(ref)
    loss = m(torch.zeros(2, 12), torch.ones(2, 16, 64))
    assert float(loss) == 1.0 and set(ref.loaded) == set(m.state_dict())


def test_precision_table_matches_the_c_header_and_defaults(monkeypatch):
    """The Python precision names map to the enum values of include/mdt_b200.h; the default mode is the fp16-operand one and the
    chunk size falls back to 4096 rows without a device (32 rows per SM with one), MDT_MAX_BATCH / MDT_PRECISION override both."""
    import re

    from moleculediffusiontransformer_b200 import _capi
    from moleculediffusiontransformer_b200 import plan as planmod

    hdr = open(os.path.join(ROOT, "include", "mdt_b200.h")).read()
    enum = {m.group(1): int(m.group(2)) for m in re.finditer(r"MDT_PREC_(\w+)\s*=\s*(\d+)", hdr)}
    assert enum == {"FP32": 0, "TF32": 1, "BF16": 2, "F16": 3}
    assert _capi.PRECISIONS == {"fp32": 0, "tf32": 1, "bf16": 2, "fp16": 3}
    monkeypatch.delenv("MDT_PRECISION", raising=False)
    monkeypatch.delenv("MDT_MAX_BATCH", raising=False)
    assert planmod.default_precision() == "fp16"
    if not torch.cuda.is_available():
        assert planmod.default_max_batch() == 4096
    monkeypatch.setenv("MDT_PRECISION", "tf32")
    monkeypatch.setenv("MDT_MAX_BATCH", "1024")
    assert planmod.default_precision() == "tf32" and planmod.default_max_batch() == 1024

"""Import the UNMODIFIED reference (lamm-mit/MoleculeDiffusionTransformer) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``moleculediffusiontransformer_b200/`` imports this.
It is used by ``oracle/make_golden.py`` (to produce ``tests/golden/*.npz``) and by the
``-m "not gpu"`` tests that pin ``oracle/unet_oracle.py`` against the real reference when
``/root/reference`` is present (it is absent on the GPU box; those tests skip there).

The reference's ``import MoleculeDiffusion`` drags in tensorflow / rdkit / seaborn /
matplotlib / torch_geometric / ipywidgets at module scope (generative.py:15-24,920-930,
transformer.py:10,17,4794-4796, diffusion.py:15) although none of them is touched by
``QMDiffusion.sample``.  We register permissive stub modules for the absent ones, import,
then remove the fake ``tensorflow`` again (einops picks its backend by scanning
``sys.modules``) and replace the notebook tqdm with a pass-through iterator.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MDT_REFERENCE_ROOT", "/root/reference")

_STUBBED = [
    "tensorflow", "tensorflow.keras", "tensorflow.keras.preprocessing",
    "tensorflow.keras.preprocessing.text", "tensorflow.keras.preprocessing.sequence",
    "seaborn", "rdkit", "rdkit.Chem", "rdkit.Chem.Draw", "rdkit.Chem.Draw.IPythonConsole",
    "rdkit.Chem.Draw.rdDepictor", "rdkit.Chem.rdFMCS", "rdkit.Chem.AllChem", "rdkit.DataStructs",
    "matplotlib", "matplotlib.pyplot", "torch_geometric", "torch_geometric.nn",
    "torch_geometric.utils", "torch_geometric.data", "ipywidgets", "torchvision",
    "torchvision.transforms", "torchvision.utils",
]


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub and which can be called."""

    def __init__(self, name: str):
        super().__init__(name)
        self.__path__ = []  # behave like a package
        self.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=True)

    def __getattr__(self, item: str):
        if item.startswith("__"):
            raise AttributeError(item)
        child = _Stub(f"{self.__name__}.{item}")
        setattr(self, item, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")

    def __mro_entries__(self, bases):  # allows `class X(stub.Something)`
        return (object,)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MoleculeDiffusion"))


_cached = None


def load():
    """Return the imported reference package (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    installed = []
    for name in _STUBBED:
        root = name.split(".")[0]
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(root) is not None and root not in {n.split(".")[0] for n in installed}:
                continue  # the real thing exists: use it
        except (ImportError, ValueError):
            pass
        sys.modules[name] = _Stub(name)
        installed.append(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    pkg = importlib.import_module("MoleculeDiffusion")
    for name in list(sys.modules):
        if name == "tensorflow" or name.startswith("tensorflow."):
            if isinstance(sys.modules[name], _Stub):
                del sys.modules[name]
    diff = importlib.import_module("MoleculeDiffusion.diffusion")
    diff.tqdm = lambda it, *a, **k: it
    _cached = pkg
    return pkg


@contextlib.contextmanager
def injected_noise(noise0, step_noise):
    """Replay recorded noise through the reference's two RNG call sites.

    ``torch.randn`` is drawn once (generative.py:853 / :163), ``torch.randn_like`` once per
    ADPM2 iteration (diffusion.py:514).  ``step_noise`` is a sequence indexed by iteration.
    """
    import torch

    real_randn, real_randn_like = torch.randn, torch.randn_like
    state = {"i": 0, "used0": False}

    def fake_randn(*shape, **kw):
        assert not state["used0"], "reference drew torch.randn more than once"
        state["used0"] = True
        assert tuple(noise0.shape) == tuple(shape if not isinstance(shape[0], (tuple, list)) else shape[0])
        return noise0.clone()

    def fake_randn_like(t, **kw):
        n = step_noise[state["i"]]
        state["i"] += 1
        assert n.shape == t.shape
        return n.clone().to(t.device)

    torch.randn, torch.randn_like = fake_randn, fake_randn_like
    try:
        yield state
    finally:
        torch.randn, torch.randn_like = real_randn, real_randn_like


def build_model(kind: str, seed: int = 0, **kw):
    """Construct a reference model with ``torch.manual_seed(seed)`` random init."""
    import io
    import torch

    load()
    gen = importlib.import_module("MoleculeDiffusion.generative")
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "inverse":
            m = gen.QMDiffusion(**kw)
        elif kind == "forward":
            m = gen.QMDiffusionForward(**kw)
        elif kind in ("analog_sparse", "analog_full"):
            gm = importlib.import_module("MoleculeDiffusion.graphmodel")
            m = (gm.AnalogDiffusionSparse if kind == "analog_sparse" else gm.AnalogDiffusionFull)(**kw)
        else:
            raise ValueError(kind)
    return m.eval()

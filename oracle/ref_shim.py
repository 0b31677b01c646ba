"""Import shim for the live reference (khangklj/Video2Music at /root/reference).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (video2music_b200/) may
import this module.  It exists so that (a) the golden vectors under
tests/golden/ can be regenerated from the *unmodified* reference
(`oracle/make_golden.py`) and (b) the CPU restatement in `oracle/amt_oracle.py`
can be validated against the reference in a container that has /root/reference
mounted.  /root/reference does not exist on the GPU box: every caller must
guard with `reference_available()`.

Why a shim is needed (all import-only, none carries hot-path arithmetic;
see SURVEY.md section 8c):
  * model/moe.py:13, model/rpr.py:11, model/custom_transformer.py:7 and
    model/grouped_query_attention.py:7 do `from torch.nn.init import *` and
    rely on it leaking `Tensor`, `math`, `warnings` -- true for the pinned
    torch 2.3.1, false for torch >= 2.4 which defines `__all__`.
  * utilities/constants.py:2 -> third_party/midi_processor/processor.py:1
    imports pretty_midi; video_music_transformer.py:17 imports gensim;
    several files import efficient_kan / lion_pytorch / seaborn / matplotlib /
    minGRU_pytorch.  None of them is installed here and none is used by the
    hot path, so empty stub modules are registered in sys.modules.
  * model/moe.py:19 and utilities/run_model_vevo.py:18 call
    parse_train_args() at import time -> sys.argv must be a bare program name.
  * generate() opens dataset/vevo_meta/*.json by relative path
    (video_music_transformer.py:1052-1057) -> cwd must be the reference root.
"""
import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("V2M_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "video_music_transformer.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _FakeWV:
    def __init__(self, vectors):
        self.vectors = vectors


class _FakeWord2Vec:
    """Stand-in for gensim.models.Word2Vec: `.load()` returns an object whose
    `.wv.vectors` is a seeded (159, 512) float32 array.  The real
    word2vec_filled.bin needs gensim, which is not installed; the array only
    initialises a frozen nn.Embedding (video_music_transformer.py:933-937)."""

    @staticmethod
    def load(path):
        import numpy as np
        rng = np.random.RandomState(20240607)
        obj = _FakeWord2Vec()
        obj.wv = _FakeWV((rng.standard_normal((159, 512)) * 0.5).astype("float32"))
        return obj


_loaded = {}


def load_reference():
    """Returns a namespace of the reference's hot-path modules."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import torch
    import torch.nn.init as _init
    if hasattr(_init, "__all__"):
        del _init.__all__
    # names that `from torch.nn.init import *` leaked on torch 2.3.1
    import math as _math
    import warnings as _warnings
    for k, v in (("Tensor", torch.Tensor), ("math", _math), ("warnings", _warnings), ("torch", torch)):
        if not hasattr(_init, k):
            setattr(_init, k, v)

    class _KANLinear(torch.nn.Module):  # import-only stub (OUT OF SCOPE: efficient_kan)
        def __init__(self, *a, **k):
            super().__init__()
            raise RuntimeError("efficient_kan is not available (out of scope)")

    _stub("pretty_midi", Note=object, PrettyMIDI=object, Instrument=object, ControlChange=object)
    g = _stub("gensim")
    gm = _stub("gensim.models", Word2Vec=_FakeWord2Vec)
    g.models = gm
    _stub("efficient_kan", KANLinear=_KANLinear)
    _stub("lion_pytorch", Lion=object)
    _stub("seaborn")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    mg = _stub("minGRU_pytorch", minGRU=object)
    mg.__path__ = []                                          # model/minGRULM.py:6 imports the submodule minGRU_pytorch.minGRU
    mg.minGRU = _stub("minGRU_pytorch.minGRU", minGRU=object)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [old_argv[0] if old_argv else "prog"]
    os.chdir(REFERENCE_ROOT)
    try:
        import importlib
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            _loaded["constants"] = importlib.import_module("utilities.constants")
            _loaded["device"] = importlib.import_module("utilities.device")
            _loaded["rpr"] = importlib.import_module("model.rpr")
            _loaded["vmt"] = importlib.import_module("model.video_music_transformer")
            _loaded["moe"] = importlib.import_module("model.moe")
            _loaded["gqa"] = importlib.import_module("model.grouped_query_attention")
            _loaded["pscan"] = importlib.import_module("model.pscan")
            _loaded["mamba"] = importlib.import_module("model.mamba")
            _loaded["custom_transformer"] = importlib.import_module("model.custom_transformer")
            _loaded["positional_encoding"] = importlib.import_module("model.positional_encoding")
            try:
                _loaded["bimamba"] = importlib.import_module("model.bimamba")
            except Exception as e:  # pragma: no cover
                _loaded["bimamba"] = None
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def reference_cwd():
    """generate() reads dataset/vevo_meta/*.json relative to cwd."""
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)

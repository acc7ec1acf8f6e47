"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference (Luckick/EAGCN) for checking.

Only tests/, tests/golden/make_golden.py, __graft_entry__ and bench.py's cpu/reference legs
may import this.  /root/reference exists in the build container only (never on the GPU box);
``available()`` says whether it can be used.

reference eagcn_pytorch/models.py:4 does ``from utils import *`` and utils.py:5,19-21,818 /
neural_fp.py:4-11 import rdkit + matplotlib at module top although the model code never
touches them; both are absent here, so inert stub modules are placed in sys.modules first
(SURVEY.md 8(c)).  The reference sources are imported from where they lie -- never copied.
"""
import importlib
import os
import sys
import types

REF_DIR = os.environ.get("EAGCN_REFERENCE_DIR", "/root/reference/eagcn_pytorch")

_STUBS = [
    "rdkit", "rdkit.Chem", "rdkit.Chem.AllChem", "rdkit.Chem.Descriptors", "rdkit.Chem.rdMolDescriptors",
    "rdkit.Chem.EState", "rdkit.Chem.rdPartialCharges", "rdkit.Chem.rdChemReactions",
    "rdkit.Chem.SaltRemover", "rdkit.Chem.Draw", "rdkit.Chem.rdmolops", "rdkit.Chem.Crippen",
    "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
]


class _Inert(types.ModuleType):
    def __getattr__(self, name):            # any attribute -> another inert stub
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Inert(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "layers.py"))


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Inert(name)


def load():
    """Return (layers, models, utils) modules of the unmodified reference."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_DIR)
    _install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    # the reference modules are called 'layers' / 'models' / 'utils'; load them under private
    # names so they can never shadow this repo's own modules
    mods = {}
    saved = {k: sys.modules.get(k) for k in ("layers", "models", "utils", "neural_fp")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        mods["layers"] = importlib.import_module("layers")
        mods["utils"] = importlib.import_module("utils")
        mods["models"] = importlib.import_module("models")
    finally:
        for k, v in saved.items():
            cur = sys.modules.pop(k, None)
            if cur is not None:
                sys.modules["_eagcn_reference_" + k] = cur
            if v is not None:
                sys.modules[k] = v
        if REF_DIR in sys.path:
            sys.path.remove(REF_DIR)
    return mods["layers"], mods["models"], mods["utils"]

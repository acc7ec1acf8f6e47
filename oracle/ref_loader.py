"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference (Luckick/EAGCN) for checking.

Only tests/, tests/golden/make_golden.py, __graft_entry__ and bench.py's cpu/reference legs
may import this.  /root/reference exists in the build container only (never on the GPU box); there the same
four modules are found in the git-ignored copy ``oracle/_ref/`` that ``tools/make_oracle_ref.sh`` produces in the
build container (it travels with the gpurun snapshot, never with git).  ``available()`` says whether either exists.

reference eagcn_pytorch/models.py:4 does ``from utils import *`` and utils.py:5,19-21,818 /
neural_fp.py:4-11 import rdkit + matplotlib at module top although the model code never
touches them; both are absent here, so inert stub modules are placed in sys.modules first
(SURVEY.md 8(c)).  The reference sources are imported from where they lie -- never copied.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("EAGCN_REFERENCE_DIR"), "/root/reference/eagcn_pytorch", os.path.join(_HERE, "_ref")]


def _pick_dir():
    for d in _CANDIDATES:
        if d and os.path.isfile(os.path.join(d, "layers.py")):
            return d
    return _CANDIDATES[1]


REF_DIR = _pick_dir()

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


def source() -> str:
    """'reference' (/root/reference itself), '_ref' (the shipped copy oracle/_ref) or 'absent'."""
    if not available():
        return "absent"
    return "_ref" if os.path.abspath(REF_DIR) == os.path.join(_HERE, "_ref") else "reference"


class cpu_only:
    """The reference picks its device from ``torch.cuda.is_available()`` at import AND at construction time
    (layers.py:10-14, :67-71, :275).  Inside this context it sees no GPU, so modules loaded / built here are the
    reference's CPU path even on the GPU box."""

    def __enter__(self):
        import torch
        self._orig = torch.cuda.is_available
        torch.cuda.is_available = lambda: False
        return self

    def __exit__(self, *exc):
        import torch
        torch.cuda.is_available = self._orig
        return False


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

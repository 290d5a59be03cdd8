"""
TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (/root/reference/ART) in this
container so that golden vectors can be generated from it.

The reference cannot be imported as-is here: numpy-quaternion, matplotlib, pyvista, pyvistaqt
and colorcet are not installed.  The hot path needs none of the plotting stack and only a
30-line piece of quaternion algebra, so:
  * `quaternion` resolves to oracle/refshim/quaternion.py,
  * the plotting modules resolve to empty stub modules.

Nothing under attosecondraytracing_b200/, bench.py's GPU arm or the `-m gpu` tests imports
this module; /root/reference does not exist on the GPU box -- there bench.py's CPU legs import the
byte-for-byte copy oracle/make_ref.py left in oracle/_ref.
"""
import contextlib
import importlib
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# where the reference package `ART` is imported from: $ART_REFERENCE_ROOT, else the tree itself in the build
# container, else the travelling byte-for-byte copy that oracle/make_ref.py leaves in oracle/_ref (GPU box)
_TRAVEL = os.path.join(os.path.dirname(_HERE), "_ref")


def _default_root():
    env = os.environ.get("ART_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/ART"):
        return "/root/reference"
    return _TRAVEL


REFERENCE_ROOT = _default_root()


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


_STUBS = [
    "matplotlib",
    "matplotlib.pyplot",
    "matplotlib.patches",
    "matplotlib.cm",
    "matplotlib.colors",
    "mpl_toolkits",
    "mpl_toolkits.mplot3d",
    "mpl_toolkits.axes_grid1",
    "pyvista",
    "pyvistaqt",
    "colorcet",
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ART"))


def load(root=None):
    """Return a namespace with the reference's hot-path modules (mp, mgeo, mmirror, ...), imported from
    `root` (default: REFERENCE_ROOT).  One root per process: `ART.*` lands in sys.modules."""
    global REFERENCE_ROOT
    if root is not None:
        REFERENCE_ROOT = root
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
                continue
            except Exception:
                pass
            mod = _StubModule(name)
            sys.modules[name] = mod
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, mod)
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)  # quaternion.py
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        ns = types.SimpleNamespace(
            mp=importlib.import_module("ART.ModuleProcessing"),
            mgeo=importlib.import_module("ART.ModuleGeometry"),
            mmirror=importlib.import_module("ART.ModuleMirror"),
            mmask=importlib.import_module("ART.ModuleMask"),
            msupp=importlib.import_module("ART.ModuleSupport"),
            mdef=importlib.import_module("ART.ModuleDefects"),
            mdet=importlib.import_module("ART.ModuleDetector"),
            moe=importlib.import_module("ART.ModuleOpticalElement"),
            moc=importlib.import_module("ART.ModuleOpticalChain"),
            mray=importlib.import_module("ART.ModuleOpticalRay"),
            msource=importlib.import_module("ART.ModuleSource"),
            mplots=importlib.import_module("ART.ModuleAnalysisAndPlots"),
        )
    return ns


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield

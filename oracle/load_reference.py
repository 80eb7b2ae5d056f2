"""Load the reference's GNN layer source UNMODIFIED from ``/root/reference``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works only in the build
container (the GPU box has no ``/root/reference``); everything that runs on the
GPU box uses the committed fixtures under ``tests/golden/`` instead.

``import neural_lam`` itself fails here (it pulls pytorch_lightning, xarray,
cartopy, ... at package import, reference ``neural_lam/__init__.py:7-14``), but
``neural_lam/gnn_layers.py`` and ``neural_lam/utils/networks.py`` only need
``torch`` and ``torch_geometric``.  They are loaded with ``importlib`` under a
synthetic parent package so that ``from . import utils`` (gnn_layers.py:10)
resolves to ``networks.py``.
"""
import importlib.util
import os
import sys
import types

from . import pyg_standin

REFERENCE_ROOT = os.environ.get("NLAM_REFERENCE_ROOT", "/root/reference")
_PKG = "_nlam_reference_src"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "neural_lam", "gnn_layers.py"))


def load():
    """Return a namespace with the reference ``InteractionNet``, ``PropagationNet``,
    ``SplitMLPs``, ``GNN_TYPES``, ``make_mlp`` loaded from the reference files."""
    if not available():
        raise FileNotFoundError(f"reference source not found under {REFERENCE_ROOT}")
    if _PKG + ".gnn_layers" in sys.modules:
        gl = sys.modules[_PKG + ".gnn_layers"]
        nw = sys.modules[_PKG + ".utils"]
    else:
        pyg_standin.install()
        parent = types.ModuleType(_PKG)
        parent.__path__ = []  # mark as package
        sys.modules[_PKG] = parent

        def _load(name, relpath):
            spec = importlib.util.spec_from_file_location(
                f"{_PKG}.{name}", os.path.join(REFERENCE_ROOT, relpath)
            )
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"{_PKG}.{name}"] = mod
            spec.loader.exec_module(mod)
            setattr(parent, name, mod)
            return mod

        nw = _load("utils", "neural_lam/utils/networks.py")
        gl = _load("gnn_layers", "neural_lam/gnn_layers.py")
    ns = types.SimpleNamespace(
        InteractionNet=gl.InteractionNet,
        PropagationNet=gl.PropagationNet,
        SplitMLPs=gl.SplitMLPs,
        GNN_TYPES=gl.GNN_TYPES,
        get_gnn_class=gl.get_gnn_class,
        make_mlp=nw.make_mlp,
        gnn_layers=gl,
        networks=nw,
    )
    return ns


def reference_test_source():
    """Text of the reference's ``tests/test_gnn_layers.py`` sections A-H (cut at
    ``# Section I``) with its ``neural_lam``/``tests`` imports stripped, for
    ``exec`` against any InteractionNet/PropagationNet implementation."""
    path = os.path.join(REFERENCE_ROOT, "tests", "test_gnn_layers.py")
    with open(path, encoding="utf-8") as f:
        lines = f.read().split("\n")
    out = []
    for line in lines:
        if "# Section I" in line:
            break
        if line.startswith("from neural_lam") or line.startswith("from tests"):
            continue
        out.append(line)
    return "\n".join(out)

"""Run the reference's graph builder ``neural_lam/create_graph.py`` UNMODIFIED (build container only).

TEST INFRASTRUCTURE ONLY.  ``create_graph`` needs networkx (present), torch_geometric's ``from_networkx`` (absent:
restated below from the PyG 2.3.1 semantics — node order = ``G.nodes()``, edge order = ``G.edges()``, node / edge
attributes stacked into tensors), matplotlib and loguru (absent, only used for plots / log lines: empty stand-ins) and
two sibling modules it imports but does not use inside ``create_graph`` itself (``.config``, ``.datastore.base``:
stand-ins).  ``reference_create_graph(xy, ...)`` returns the tensors the reference writes to its graph directory.
"""
import importlib.util
import os
import sys
import tempfile
import types
from collections import defaultdict

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("NLAM_REFERENCE_ROOT", "/root/reference")
_PKG = "_nlam_reference_cg"


def available():
    try:
        import networkx  # noqa: F401
    except Exception:
        return False
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "neural_lam", "create_graph.py"))


class _Data:
    """Attribute bag standing in for ``torch_geometric.data.Data``."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def clone(self):
        return _Data(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()})


def from_networkx(G):
    """PyG 2.3.1 ``torch_geometric.utils.convert.from_networkx`` restated for the graphs ``create_graph`` builds."""
    import networkx as nx

    G = G.to_directed() if not nx.is_directed(G) else G
    mapping = dict(zip(G.nodes(), range(G.number_of_nodes())))
    edge_index = torch.empty((2, G.number_of_edges()), dtype=torch.long)
    for i, (src, dst) in enumerate(G.edges()):
        edge_index[0, i] = mapping[src]
        edge_index[1, i] = mapping[dst]
    data = defaultdict(list)
    node_attrs = list(next(iter(G.nodes(data=True)))[-1].keys()) if G.number_of_nodes() > 0 else []
    for _, feat in G.nodes(data=True):
        for k, v in feat.items():
            data[str(k)].append(v)
    for _, _, feat in G.edges(data=True):
        for k, v in feat.items():
            data[str(f"edge_{k}" if k in node_attrs else k)].append(v)
    out = {}
    for k, v in data.items():
        try:
            out[k] = torch.stack(v) if isinstance(v[0], torch.Tensor) else torch.as_tensor(np.array(v))
        except Exception:
            out[k] = v
    out["edge_index"] = edge_index.view(2, -1)
    return _Data(**out)


def _install_standins():
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = mod("matplotlib")
            mpl.__path__ = []
            mod("matplotlib.pyplot")
            mpl.pyplot = sys.modules["matplotlib.pyplot"]
            mpl.figure = types.SimpleNamespace(Figure=object)   # only named in a return annotation of plot_graph
            mpl.axes = types.SimpleNamespace(Axes=object)
    if "loguru" not in sys.modules:
        try:
            import loguru  # noqa: F401
        except Exception:
            mod("loguru", logger=types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None,
                                                       debug=lambda *a, **k: None, error=lambda *a, **k: None))
    from . import pyg_standin

    pyg_standin.install()
    pyg = sys.modules["torch_geometric"]
    if not hasattr(pyg, "__path__"):
        pyg.__path__ = []
    mod("torch_geometric.data", Data=_Data)
    pyg.data = sys.modules["torch_geometric.data"]
    utils = mod("torch_geometric.utils")
    utils.__path__ = []
    pyg.utils = utils
    conv = mod("torch_geometric.utils.convert", from_networkx=from_networkx)
    utils.convert = conv


def load():
    if f"{_PKG}.create_graph" in sys.modules:
        return sys.modules[f"{_PKG}.create_graph"]
    _install_standins()
    parent = types.ModuleType(_PKG)
    parent.__path__ = []
    sys.modules[_PKG] = parent
    cfg = types.ModuleType(f"{_PKG}.config")
    cfg.load_config_and_datastore = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stand-in"))
    sys.modules[f"{_PKG}.config"] = cfg
    ds = types.ModuleType(f"{_PKG}.datastore")
    ds.__path__ = []
    sys.modules[f"{_PKG}.datastore"] = ds
    base = types.ModuleType(f"{_PKG}.datastore.base")
    base.BaseRegularGridDatastore = type("BaseRegularGridDatastore", (), {})
    sys.modules[f"{_PKG}.datastore.base"] = base
    spec = importlib.util.spec_from_file_location(f"{_PKG}.create_graph",
                                                  os.path.join(REFERENCE_ROOT, "neural_lam", "create_graph.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[f"{_PKG}.create_graph"] = m
    spec.loader.exec_module(m)
    return m


def reference_create_graph(Nx, Ny, hierarchical=False, n_max_levels=None, spacing=1.0):
    """{file stem: tensor / list of tensors} written by the reference's ``create_graph`` for a regular Nx x Ny grid with
    coordinates (i * spacing, j * spacing), i.e. the grid ``neural_lam_b200.synthetic.make_graph_spec`` assumes."""
    cg = load()
    gx, gy = np.meshgrid(np.arange(Nx) * spacing, np.arange(Ny) * spacing, indexing="ij")
    xy = np.stack([gx, gy], axis=-1).astype(np.float64)  # (Nx, Ny, 2)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        cg.create_graph(d, xy, n_max_levels=n_max_levels, hierarchical=hierarchical, create_plot=False)
        for fn in os.listdir(d):
            if fn.endswith(".pt"):
                out[fn[:-3]] = torch.load(os.path.join(d, fn), weights_only=True)
    return out

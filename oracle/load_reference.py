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
import re
import sys
import types

from . import pyg_standin

REFERENCE_ROOT = os.environ.get("NLAM_REFERENCE_ROOT", "/root/reference")
_PKG = "_nlam_reference_src"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "neural_lam", "gnn_layers.py"))


def _load_file(name, relpath):
    """exec one reference source file, unmodified, as module ``_PKG.name``."""
    full = f"{_PKG}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    parent_name, _, leaf = full.rpartition(".")
    setattr(sys.modules[parent_name], leaf, mod)
    return mod


def _synthetic_package(name):
    full = f"{_PKG}.{name}" if name else _PKG
    if full in sys.modules:
        return sys.modules[full]
    mod = types.ModuleType(full)
    mod.__path__ = []  # mark as package
    sys.modules[full] = mod
    if name:
        parent_name, _, leaf = full.rpartition(".")
        setattr(sys.modules[parent_name], leaf, mod)
    return mod


def load():
    """Return a namespace with the reference ``InteractionNet``, ``PropagationNet``,
    ``SplitMLPs``, ``GNN_TYPES``, ``make_mlp`` loaded from the reference files."""
    if not available():
        raise FileNotFoundError(f"reference source not found under {REFERENCE_ROOT}")
    pyg_standin.install()
    _synthetic_package("")
    # ``neural_lam/utils/__init__.py`` pulls logging / plotting helpers (matplotlib, ...): the package
    # object is synthetic and re-exports what the loaded files define, the FILES are the reference's
    utils = _synthetic_package("utils")
    nw = _load_file("utils.networks", "neural_lam/utils/networks.py")
    utils.make_mlp, utils.make_gnn_seq = nw.make_mlp, nw.make_gnn_seq
    gl = _load_file("gnn_layers", "neural_lam/gnn_layers.py")
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


class _Values:
    """The ``.values`` face of the xarray objects the step predictors read."""

    def __init__(self, arr):
        self.values = arr


class StubDatastore:
    """The ``BaseDatastore`` surface the reference's step predictors / ARForecaster touch
    (step_predictors/base.py:62-106, :196; graph/base.py:86-135; utils/graph.py:443-446, :494-512;
    forecasters/autoregressive.py:35-40), filled from a ``neural_lam_b200.synthetic.SyntheticDatastore``;
    the graph is read by the reference's own ``load_graph`` from ``root_path/graph/<name>``."""

    def __init__(self, ds, root_path):
        import pathlib

        self._ds = ds
        self.root_path = pathlib.Path(root_path)
        self.num_grid_points = ds.num_grid_nodes
        self.boundary_mask = _Values(ds.boundary_mask.reshape(-1).numpy())

    def get_num_data_vars(self, category):
        return {"state": self._ds.num_state_vars, "forcing": self._ds.num_forcing_vars,
                "static": self._ds.num_static_vars}[category]

    def get_vars_names(self, category):
        assert category == "state"
        return list(self._ds.state_var_names)

    def get_dataarray(self, category, split=None, standardize=False):
        assert category == "static"
        if self._ds.num_static_vars == 0:
            return None
        return _Values(self._ds.grid_static_features.numpy())

    def get_standardization_dataarray(self, category):
        assert category == "state"
        d = self._ds
        return types.SimpleNamespace(
            state_mean=_Values(d.state_mean.numpy()), state_std=_Values(d.state_std.numpy()),
            state_diff_mean_standardized=_Values(d.state_diff_mean.numpy()),
            state_diff_std_standardized=_Values(d.state_diff_std.numpy()))

    def get_xy_extent(self, category):
        gxy = self._ds.grid_xy
        return [float(gxy[:, 0].min()), float(gxy[:, 0].max()), float(gxy[:, 1].min()), float(gxy[:, 1].max())]


def load_models():
    """Load the reference's step predictors and forecaster UNMODIFIED:
    ``models/step_predictors/base.py``, ``graph/{base,graph_lam,hierarchical,hi_lam,hi_lam_parallel}.py``,
    ``models/forecasters/{base,autoregressive}.py``, ``utils/{graph,buffer_list,tensor}.py`` — under the
    synthetic parent package, with stubs only for what those files import but never compute with here:
    ``datastore.BaseDatastore`` (a type annotation), ``create_graph``'s two format constants (read from the
    reference file), ``utils.log_on_rank_zero`` (a print).  Returns a namespace with the classes."""
    ns = load()
    pkg = sys.modules[_PKG]
    utils = sys.modules[f"{_PKG}.utils"]
    if not hasattr(pkg, "datastore"):
        dsm = _synthetic_package("datastore")

        class BaseDatastore:  # annotation only on the path
            pass

        dsm.BaseDatastore = BaseDatastore
        cg = _synthetic_package("create_graph")
        text = open(os.path.join(REFERENCE_ROOT, "neural_lam", "create_graph.py"), encoding="utf-8").read()
        for const in ("METAINFO_FILENAME", "CURRENT_GRAPH_SPEC_VERSION"):
            m = re.search(rf'^{const}\s*=\s*"([^"]+)"', text, re.M)
            setattr(cg, const, m.group(1))
    bl = _load_file("utils.buffer_list", "neural_lam/utils/buffer_list.py")
    tn = _load_file("utils.tensor", "neural_lam/utils/tensor.py")
    gr = _load_file("utils.graph", "neural_lam/utils/graph.py")
    utils.BufferList = bl.BufferList
    utils.inverse_sigmoid, utils.inverse_softplus = tn.inverse_sigmoid, tn.inverse_softplus
    utils.load_graph, utils.load_and_register_graph = gr.load_graph, gr.load_and_register_graph
    utils.compute_grid_input_dim = gr.compute_grid_input_dim
    utils.log_on_rank_zero = lambda *a, **k: None
    for sub in ("models", "models.step_predictors", "models.step_predictors.graph", "models.forecasters"):
        _synthetic_package(sub)
    sp = _load_file("models.step_predictors.base", "neural_lam/models/step_predictors/base.py")
    gdir = "neural_lam/models/step_predictors/graph/"
    gb = _load_file("models.step_predictors.graph.base", gdir + "base.py")
    glam = _load_file("models.step_predictors.graph.graph_lam", gdir + "graph_lam.py")
    hier = _load_file("models.step_predictors.graph.hierarchical", gdir + "hierarchical.py")
    hil = _load_file("models.step_predictors.graph.hi_lam", gdir + "hi_lam.py")
    hip = _load_file("models.step_predictors.graph.hi_lam_parallel", gdir + "hi_lam_parallel.py")
    _load_file("models.forecasters.base", "neural_lam/models/forecasters/base.py")
    ar = _load_file("models.forecasters.autoregressive", "neural_lam/models/forecasters/autoregressive.py")
    ns.StepPredictor, ns.BaseGraphModel, ns.GraphLAM = sp.StepPredictor, gb.BaseGraphModel, glam.GraphLAM
    ns.BaseHiGraphModel, ns.HiLAM, ns.HiLAMParallel = hier.BaseHiGraphModel, hil.HiLAM, hip.HiLAMParallel
    ns.ARForecaster = ar.ARForecaster
    ns.load_graph = gr.load_graph
    ns.graph_utils = gr
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

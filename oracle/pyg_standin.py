"""Minimal stand-in for ``torch_geometric.nn.{MessagePassing, Sequential}``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference pins ``torch-geometric == 2.3.1`` (reference ``pyproject.toml:33``)
without ``torch-scatter``/``pyg-lib``; that package is not installable in this
image (no network, not in the wheelhouse).  The aggregation arithmetic of the
hot path lives there, so it is restated here from the published semantics of that
version, anchored on the reference's call sites:

* subclassing / ``super().__init__(aggr=aggr)``      reference gnn_layers.py:14, :67
* ``self.propagate(edge_index, x=..., edge_attr=...)`` reference gnn_layers.py:145-147
* ``super().aggregate(inputs, index, ptr, dim_size)``  reference gnn_layers.py:188
* ``pyg.nn.Sequential("mesh_rep, edge_rep", [...])``   reference networks.py:93,
  graph_lam.py:117

PyG 2.3.1 semantics restated (flow="source_to_target", node_dim=-2):
``x_j = x.index_select(-2, edge_index[0])``, ``x_i = x.index_select(-2,
edge_index[1])``; sum aggregation = ``zeros(dim_size).scatter_add_(-2,
broadcast(index), src)``; mean = that divided by the per-receiver count clamped
to >= 1.

This is enough to run the reference's ``gnn_layers.py`` / ``networks.py``
unmodified (see ``oracle/load_reference.py``).
"""
import sys
import types

import torch


class MessagePassing(torch.nn.Module):
    """Restatement of the part of PyG's MessagePassing the reference uses."""

    def __init__(self, aggr="sum", flow="source_to_target", node_dim=-2):
        super().__init__()
        if aggr == "add":
            aggr = "sum"
        assert aggr in ("sum", "mean")
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs["x"]
        x_j = x.index_select(self.node_dim, edge_index[0])
        x_i = x.index_select(self.node_dim, edge_index[1])
        msg = self.message(x_j=x_j, x_i=x_i, edge_attr=kwargs["edge_attr"])
        return self.aggregate(msg, edge_index[1], None, x.size(self.node_dim))

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        dim = self.node_dim if self.node_dim >= 0 else inputs.dim() + self.node_dim
        dim_size = int(dim_size)
        size = list(inputs.shape)
        size[dim] = dim_size
        view = [1] * inputs.dim()
        view[dim] = -1
        idx = index.view(view).expand_as(inputs)
        out = inputs.new_zeros(size).scatter_add_(dim, idx, inputs)
        if self.aggr == "mean":
            count = inputs.new_zeros(dim_size).scatter_add_(
                0, index, inputs.new_ones(index.numel())
            )
            out = out / count.clamp(min=1).view(view)
        return out

    def message(self, **kwargs):  # pragma: no cover - always overridden
        raise NotImplementedError


class Sequential(torch.nn.Module):
    """Restatement of ``pyg.nn.Sequential`` for the one signature the reference
    uses: ``"mesh_rep, edge_rep"`` with modules
    ``"mesh_rep, mesh_rep, edge_rep -> mesh_rep, edge_rep"``.  Children are
    registered as ``module_{i}`` like PyG 2.3 does."""

    def __init__(self, input_args, modules):
        super().__init__()
        assert input_args.replace(" ", "") == "mesh_rep,edge_rep"
        self._n = len(modules)
        for i, (mod, desc) in enumerate(modules):
            assert (
                desc.replace(" ", "")
                == "mesh_rep,mesh_rep,edge_rep->mesh_rep,edge_rep"
            )
            self.add_module(f"module_{i}", mod)

    def forward(self, mesh_rep, edge_rep):
        for i in range(self._n):
            mesh_rep, edge_rep = getattr(self, f"module_{i}")(
                mesh_rep, mesh_rep, edge_rep
            )
        return mesh_rep, edge_rep


def install():
    """Register the stand-in as ``torch_geometric`` in ``sys.modules`` unless a
    real torch_geometric is importable."""
    try:  # pragma: no cover - real PyG absent in this image
        import torch_geometric  # noqa: F401

        return False
    except Exception:
        pass
    pkg = types.ModuleType("torch_geometric")
    nn_mod = types.ModuleType("torch_geometric.nn")
    nn_mod.MessagePassing = MessagePassing
    nn_mod.Sequential = Sequential
    pkg.nn = nn_mod
    pkg.__version__ = "2.3.1-standin"
    sys.modules["torch_geometric"] = pkg
    sys.modules["torch_geometric.nn"] = nn_mod
    return True

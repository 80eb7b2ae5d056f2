"""Host-side (CPU) logic of the drop-in layer: reference-compatible structure, index tables."""
import numpy as np
import pytest
import torch

import neural_lam_b200 as nlb
from neural_lam_b200 import InteractionNet, PropagationNet


def _rand_ei(ns, nr, ne, seed=0):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, ns, (ne,), generator=g), torch.randint(0, nr, (ne,), generator=g)])
    return ei


def test_structure_like_reference():
    # reference tests/test_gnn_layers.py section A (own restatement)
    assert issubclass(PropagationNet, InteractionNet)
    ei = _rand_ei(5, 4, 10)
    assert PropagationNet(ei, 8, aggr="sum").aggr == "mean"
    assert InteractionNet(ei.clone(), 8, aggr="sum").aggr == "sum"
    assert InteractionNet(ei.clone(), 8, aggr="mean").aggr == "mean"
    with pytest.raises(ValueError):
        InteractionNet(ei, 8, aggr="max")
    p = PropagationNet(ei, 16)
    assert p.edge_mlp[0].in_features == 48 and p.aggr_mlp[0].in_features == 32
    st = p.edge_index
    assert st[1].min() >= 0 and st[1].max() < p.num_rec and st[0].min() >= p.num_rec
    assert "edge_index" not in p.state_dict()
    assert sorted(p.state_dict()) == sorted(
        [f"{m}.{i}.{w}" for m in ("edge_mlp", "aggr_mlp") for i in (0, 2, 3) for w in ("weight", "bias")]
    )
    assert p.state_dict()["edge_mlp.0.weight"].shape == (16, 48)
    assert p.state_dict()["aggr_mlp.0.weight"].shape == (16, 32)


def test_registry():
    assert set(nlb.GNN_TYPES) == {"InteractionNet", "PropagationNet"}
    assert nlb.get_gnn_class("PropagationNet") is PropagationNet
    with pytest.raises(ValueError):
        nlb.get_gnn_class("nope")
    with pytest.raises(ValueError):
        nlb.make_gnn_seq(_rand_ei(4, 4, 8), 0, 1, 8)
    seq = nlb.make_gnn_seq(_rand_ei(4, 4, 8), 2, 1, 8)
    assert [n for n, _ in seq.named_children()] == ["module_0", "module_1"]


def test_split_mlps_state_dict_names():
    net = InteractionNet(_rand_ei(6, 4, 12), 8, edge_chunk_sizes=[5, 7], aggr_chunk_sizes=[2, 2])
    keys = list(net.state_dict())
    assert "edge_mlp.mlps.1.0.weight" in keys and "aggr_mlp.mlps.0.3.bias" in keys


@pytest.mark.parametrize("seed", range(4))
def test_index_tables(seed):
    ns, nr, ne = 13, 7, 50
    ei = _rand_ei(ns, nr, ne, seed)
    net = InteractionNet(ei, 4)
    snd, rcv = ei[0].numpy(), ei[1].numpy()
    perm = net._perm32.numpy()
    assert np.array_equal(perm, np.argsort(rcv, kind="stable"))
    assert np.array_equal(net._inv_perm32.numpy()[perm], np.arange(ne))
    rp = net._rowptr32.numpy()
    assert rp[0] == 0 and rp[-1] == ne
    for r in range(net.num_rec):
        assert np.all(rcv[perm[rp[r]:rp[r + 1]]] == r)
    sp, so = net._sptr32.numpy(), net._sorder32.numpy()
    for s in range(net.num_send_min):
        assert np.all(snd[so[sp[s]:sp[s + 1]]] == s)
    assert net.max_in_degree == np.bincount(rcv).max()


def test_synthetic_graph_degrees_match_the_kernel_selection():
    """The receiver-tiled (ELL) kernels are selected for edge sets with a uniform in-degree: mesh->grid gives every
    grid node exactly its 4 nearest mesh nodes (reference create_graph.py:779-792), hierarchical down edges give every
    node one parent; grid->mesh and the mesh graph itself have varying in-degrees (general CSR kernels)."""
    import torch
    from neural_lam_b200 import synthetic

    spec = synthetic.make_graph_spec(30, 27)
    G = 30 * 27
    deg = torch.bincount(spec["m2g_edge_index"][1], minlength=G)
    assert deg.min().item() == deg.max().item() == 4
    assert torch.bincount(spec["g2m_edge_index"][1]).unique().numel() > 1
    assert torch.bincount(spec["m2m_edge_index"][1]).unique().numel() > 1
    hspec = synthetic.make_graph_spec(30, 27, hierarchical=True)
    for ei in hspec["mesh_down_edge_index"]:
        d = torch.bincount(ei[1])
        assert d.min().item() == d.max().item() == 1


@pytest.mark.parametrize("hier", [False, True])
def test_graph_csr_cache_roundtrip(tmp_path, hier):
    """Cached CSR next to the reference-format graph files: loading applies the stored permutation, the
    result equals sorting at load time, and a stale cache is ignored."""
    import os

    import torch
    from neural_lam_b200 import synthetic
    from neural_lam_b200.models import _sort_edges

    spec = synthetic.make_graph_spec(30, 27, hierarchical=hier)
    assert spec["hierarchical"] == hier
    # shuffle one edge set so that the stored order is NOT receiver-sorted
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(spec["g2m_edge_index"].shape[1], generator=g)
    spec["g2m_edge_index"], spec["g2m_features"] = spec["g2m_edge_index"][:, perm], spec["g2m_features"][perm]
    d = str(tmp_path / "graph")
    synthetic.save_graph(spec, d)
    raw = synthetic.load_graph(d, spec["grid_xy"])
    assert torch.equal(raw["g2m_edge_index"], spec["g2m_edge_index"])  # no cache yet: stored order
    synthetic.save_graph_csr_cache(d)
    got = synthetic.load_graph(d, spec["grid_xy"])
    for k in ("g2m", "m2g"):
        ei, f = _sort_edges(raw[f"{k}_edge_index"], raw[f"{k}_features"])
        assert torch.equal(got[f"{k}_edge_index"], ei) and torch.equal(got[f"{k}_features"], f)
        assert bool((got[f"{k}_edge_index"][1][1:] >= got[f"{k}_edge_index"][1][:-1]).all())
    m2m_raw = raw["m2m_edge_index"] if hier else [raw["m2m_edge_index"]]
    m2m_got = got["m2m_edge_index"] if hier else [got["m2m_edge_index"]]
    for a, b in zip(m2m_raw, m2m_got):
        assert torch.equal(b, a[:, torch.sort(a[1], stable=True).indices])
    cache = torch.load(os.path.join(d, "g2m_csr.pt"), weights_only=True)
    assert cache["rowptr"][-1].item() == cache["n_edges"] == raw["g2m_edge_index"].shape[1]
    # stale cache (edge count changed): ignored
    torch.save(raw["g2m_edge_index"][:, :-1], os.path.join(d, "g2m_edge_index.pt"))
    torch.save(raw["g2m_features"][:-1], os.path.join(d, "g2m_features.pt"))
    stale = synthetic.load_graph(d, spec["grid_xy"])
    assert torch.equal(stale["g2m_edge_index"], raw["g2m_edge_index"][:, :-1])


def test_boundary_blocks_cover_exactly_the_masked_rows():
    """rollout_from_host moves only the boundary-mask rows of the boundary state: a frame mask becomes two runs (top, bottom) and
    one block of equally spaced runs (the strips left and right of the interior)."""
    import numpy as np
    from neural_lam_b200.models import _boundary_blocks

    for nx, ny, w in ((268, 238, 10), (30, 27, 2), (16, 16, 2)):
        m = torch.zeros(nx, ny)
        m[:w] = 1
        m[-w:] = 1
        m[:, :w] = 1
        m[:, -w:] = 1
        blocks = _boundary_blocks(m.reshape(-1, 1))
        assert blocks is not None and len(blocks) == 3
        hit = np.zeros(nx * ny)
        for start, rlen, pitch, count in blocks:
            for j in range(count):
                hit[start + j * pitch:start + j * pitch + rlen] += 1
            if count > 1:
                assert (nx * ny) % pitch == 0 and rlen <= pitch
        assert (hit == m.reshape(-1).numpy()).all()
    assert _boundary_blocks(torch.ones(64, 1)) is None                      # nothing to save
    assert _boundary_blocks(torch.zeros(64, 1)) is None
    assert _boundary_blocks((torch.arange(4000) % 3 == 0).float().reshape(-1, 1)) is None  # too fragmented

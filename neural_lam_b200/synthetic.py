"""Synthetic MEPS-shaped graphs and a minimal datastore stand-in.

The reference builds its graphs offline with networkx loops (reference
neural_lam/create_graph.py:356-862) — far too slow for >=1M-node grids and it needs
torch_geometric + matplotlib at import.  ``make_graph_spec`` is a vectorised numpy/scipy
generator that follows the same topology rules on a regular grid:

  * mesh levels: ``nx=3``, ``nlev=int(log(max(Nx,Ny))/log(3))``, ``nleaf=3**nlev``, level
    ``lev`` has ``n=nleaf/3**lev`` nodes per side (create_graph.py:438-453), positions at cell
    centres ``linspace(min+d/2, max-d/2, n)`` (:296-303), directed 8-neighbour edges (:306-329)
    with features ``[len, pos[sender]-pos[receiver]]`` (:135-138, :320-327);
  * multiscale: coarser levels are merged onto the finest level's nodes at the centre of each
    3x3 block (:596-611); hierarchical: levels stay separate, up edges = 1-nearest coarser node
    (:485-509), down edges = reversed with negated offsets (:562-569);
  * g2m: grid nodes within ``0.67*dm`` of a bottom mesh node, ``dm`` = mesh spacing along the
    1st axis (:698-758); m2g: 4 nearest bottom mesh nodes of every grid node (:779-792);
  * node order: mesh node (i,j) -> i*n+j, grid node (i,j) -> i*Ny+j (sorted labels, :667-676).

``normalize_graph`` applies what ``load_graph`` does at load time (reference
neural_lam/utils/graph.py:291-303 mesh coords / max grid span, :343-350 edge features /
longest m2m edge).  Edge ORDER within a set is free in the on-disk format; every edge set is
stored receiver-sorted here so that the kernels' CSR order is the storage order.
"""
import math
import os

import numpy as np
import torch

try:
    from scipy.spatial import cKDTree
except Exception:  # pragma: no cover
    cKDTree = None


def _level_positions(xy_min, xy_max, n):
    d = (xy_max - xy_min) / n
    lx = np.linspace(xy_min[0] + d[0] / 2, xy_max[0] - d[0] / 2, n)
    ly = np.linspace(xy_min[1] + d[1] / 2, xy_max[1] - d[1] / 2, n)
    gx, gy = np.meshgrid(lx, ly, indexing="ij")
    return np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1)  # node (i,j) -> i*n+j


def _grid8_edges(n):
    """Directed 8-neighbour edges of an n x n lattice, node (i,j) -> i*n+j."""
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    snd, rcv = [], []
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            if di == 0 and dj == 0:
                continue
            ok = (ii + di >= 0) & (ii + di < n) & (jj + dj >= 0) & (jj + dj < n)
            snd.append((ii[ok] * n + jj[ok]).reshape(-1))
            rcv.append(((ii[ok] + di) * n + (jj[ok] + dj)).reshape(-1))
    return np.concatenate(snd), np.concatenate(rcv)


def _edge_features(pos_s, pos_r):
    vd = pos_s - pos_r
    ln = np.sqrt((vd ** 2).sum(axis=1, keepdims=True))
    return np.concatenate([ln, vd], axis=1).astype(np.float32)


def _sort_by_receiver(snd, rcv, feat):
    order = np.lexsort((snd, rcv))  # receiver-major, sender-minor: deterministic
    return snd[order], rcv[order], feat[order]


def _pack(snd, rcv, feat):
    snd, rcv, feat = _sort_by_receiver(snd, rcv, feat)
    return torch.from_numpy(np.stack([snd, rcv]).astype(np.int64)), torch.from_numpy(feat)


def make_graph_spec(Nx, Ny, hierarchical=False, n_levels=None, spacing=1.0):
    """Graph tensors (UNNORMALISED, like the files ``create_graph`` writes) for a regular
    ``Nx x Ny`` grid with unit spacing.  Returns a dict with the keys ``load_graph`` produces
    plus ``grid_xy`` (G,2), ``grid_shape`` and ``hierarchical``."""
    if cKDTree is None:  # pragma: no cover
        raise RuntimeError("scipy is required for the synthetic graph generator")
    gx, gy = np.meshgrid(np.arange(Nx) * spacing, np.arange(Ny) * spacing, indexing="ij")
    grid_xy = np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1).astype(np.float64)  # (i,j) -> i*Ny+j
    xy_min = grid_xy.min(axis=0)
    xy_max = grid_xy.max(axis=0)

    nlev = int(math.log(max(Nx, Ny)) / math.log(3))
    nleaf = 3 ** nlev
    mesh_levels = nlev - 1
    if n_levels:
        mesh_levels = min(mesh_levels, n_levels)
    if mesh_levels < 1:
        raise ValueError(f"grid {Nx}x{Ny} too small for a mesh")
    ns = [nleaf // (3 ** lev) for lev in range(1, mesh_levels + 1)]
    pos = [_level_positions(xy_min, xy_max, n) for n in ns]

    spec = {"hierarchical": bool(hierarchical and mesh_levels > 1), "grid_shape": (Nx, Ny),
            "grid_xy": torch.from_numpy(grid_xy.astype(np.float32))}
    if spec["hierarchical"]:
        m2m_ei, m2m_f = [], []
        for n, p in zip(ns, pos):
            s, r = _grid8_edges(n)
            ei, f = _pack(s, r, _edge_features(p[s], p[r]))
            m2m_ei.append(ei)
            m2m_f.append(f)
        up_ei, up_f, down_ei, down_f = [], [], [], []
        for l in range(mesh_levels - 1):
            tree = cKDTree(pos[l + 1])
            nearest = tree.query(pos[l], 1)[1]
            s = np.arange(pos[l].shape[0])
            ei, f = _pack(s, nearest, _edge_features(pos[l][s], pos[l + 1][nearest]))
            up_ei.append(ei)
            up_f.append(f)
            ei, f = _pack(nearest, s, _edge_features(pos[l + 1][nearest], pos[l][s]))
            down_ei.append(ei)
            down_f.append(f)
        spec.update(m2m_edge_index=m2m_ei, m2m_features=m2m_f,
                    mesh_up_edge_index=up_ei, mesh_up_features=up_f,
                    mesh_down_edge_index=down_ei, mesh_down_features=down_f,
                    mesh_static_features=[torch.from_numpy(p.astype(np.float32)) for p in pos])
    else:
        n0 = ns[0]
        snd_all, rcv_all, feat_all = [], [], []
        for k, (n, p) in enumerate(zip(ns, pos)):
            s, r = _grid8_edges(n)
            f = _edge_features(p[s], p[r])
            # level-k node (a,b) sits on bottom node (3^k a + (3^k-1)/2, same for b)
            scale = 3 ** k
            off = (scale - 1) // 2

            def to_bottom(idx, n=n, scale=scale, off=off):
                a, b = idx // n, idx % n
                return (a * scale + off) * n0 + (b * scale + off)

            snd_all.append(to_bottom(s))
            rcv_all.append(to_bottom(r))
            feat_all.append(f)
        ei, f = _pack(np.concatenate(snd_all), np.concatenate(rcv_all), np.concatenate(feat_all))
        spec.update(m2m_edge_index=ei, m2m_features=f,
                    mesh_up_edge_index=[], mesh_up_features=[], mesh_down_edge_index=[], mesh_down_features=[],
                    mesh_static_features=torch.from_numpy(pos[0].astype(np.float32)))

    # grid <-> bottom mesh
    bottom = pos[0]
    n0 = ns[0]
    # mesh nodes (1, 0) and (0, 0): the spacing along the FIRST axis (create_graph.py:702-704 indexes the node labels
    # (0, 1, 0) and (0, 0, 0) = (level prefix, i, j))
    dm = np.sqrt(((bottom[n0] - bottom[0]) ** 2).sum())
    gtree = cKDTree(grid_xy)
    neigh = gtree.query_ball_point(bottom, dm * 0.67)
    counts = np.fromiter((len(x) for x in neigh), dtype=np.int64, count=len(neigh))
    g_idx = np.fromiter((i for x in neigh for i in x), dtype=np.int64, count=int(counts.sum()))
    m_idx = np.repeat(np.arange(bottom.shape[0]), counts)
    spec["g2m_edge_index"], spec["g2m_features"] = _pack(g_idx, m_idx, _edge_features(grid_xy[g_idx], bottom[m_idx]))
    mtree = cKDTree(bottom)
    nn = mtree.query(grid_xy, 4)[1]  # (G,4)
    g_rep = np.repeat(np.arange(grid_xy.shape[0]), 4)
    m_flat = nn.reshape(-1)
    spec["m2g_edge_index"], spec["m2g_features"] = _pack(m_flat, g_rep, _edge_features(bottom[m_flat], grid_xy[g_rep]))
    return spec


GRAPH_SPEC_VERSION = "0.1.0"  # reference create_graph.py:25 CURRENT_GRAPH_SPEC_VERSION; file name :24
METAINFO_FILENAME = "metainfo.yaml"


def normalize_graph(spec):
    """What ``load_graph`` does to the on-disk tensors (reference utils/graph.py:291-303,
    :343-350): mesh coordinates / max grid span, edge features / longest m2m edge."""
    out = dict(spec)
    gxy = spec["grid_xy"]
    span = float(max(gxy[:, 0].max() - gxy[:, 0].min(), gxy[:, 1].max() - gxy[:, 1].min()))
    span = span if span != 0 else 1.0
    hier = spec["hierarchical"]
    m2m_f = spec["m2m_features"] if hier else [spec["m2m_features"]]
    # bit-for-bit what the reference's load_graph computes (utils/graph.py:343-350, :390-394): the per-level sets
    # (BufferList ``/=``, buffer_list.py:93) are MULTIPLIED by the float32 reciprocal of the longest m2m edge, the
    # g2m / m2g features are divided by it
    longest = max(torch.max(f[:, 0]) for f in m2m_f)
    inv_longest = 1.0 / longest
    def scale_xy(m):
        # only the two coordinate columns are scaled (utils/graph.py:302-303 ``m[:, :2] /= scaling``); legacy
        # graphs store normalised mesh coordinates already (:284-289)
        if spec.get("legacy"):
            return m
        m = m.clone()
        m[:, :2] /= span
        return m

    if hier:
        out["mesh_static_features"] = [scale_xy(m) for m in spec["mesh_static_features"]]
        for k in ("m2m_features", "mesh_up_features", "mesh_down_features"):
            out[k] = [f * inv_longest for f in spec[k]]
    else:
        out["mesh_static_features"] = scale_xy(spec["mesh_static_features"])
        out["m2m_features"] = spec["m2m_features"] * inv_longest
    out["g2m_features"] = spec["g2m_features"] / longest
    out["m2g_features"] = spec["m2g_features"] / longest
    out["normalized"] = True
    return out


def save_graph(spec, graph_dir):
    """Write the reference's on-disk format (docs/graph_storage_spec.md; file set as in
    create_graph.py:366-410): lists per level for m2m / mesh / up / down."""
    os.makedirs(graph_dir, exist_ok=True)
    hier = spec["hierarchical"]
    as_list = (lambda v: v) if hier else (lambda v: [v])
    torch.save(as_list(spec["m2m_edge_index"]), os.path.join(graph_dir, "m2m_edge_index.pt"))
    torch.save(as_list(spec["m2m_features"]), os.path.join(graph_dir, "m2m_features.pt"))
    torch.save(as_list(spec["mesh_static_features"]), os.path.join(graph_dir, "mesh_features.pt"))
    for k in ("g2m", "m2g"):
        torch.save(spec[f"{k}_edge_index"], os.path.join(graph_dir, f"{k}_edge_index.pt"))
        torch.save(spec[f"{k}_features"], os.path.join(graph_dir, f"{k}_features.pt"))
    if hier:
        for k in ("mesh_up", "mesh_down"):
            torch.save(spec[f"{k}_edge_index"], os.path.join(graph_dir, f"{k}_edge_index.pt"))
            torch.save(spec[f"{k}_features"], os.path.join(graph_dir, f"{k}_features.pt"))
    with open(os.path.join(graph_dir, "metainfo.yaml"), "w", encoding="utf-8") as f:
        f.write(f"spec_version: {GRAPH_SPEC_VERSION}\n")


def load_graph(graph_dir, grid_xy, use_csr_cache=True):
    """Read a graph directory in the reference's current on-disk format (the file set
    ``load_graph`` reads, reference utils/graph.py:146-422) into an unnormalised spec.  If ``save_graph_csr_cache`` has
    been run on the directory, every edge set (and its static features) is returned receiver-sorted."""
    spec = _load_graph_raw(graph_dir, grid_xy)
    if use_csr_cache:
        for k in _EDGE_SETS:
            if f"{k}_edge_index" in spec and (not isinstance(spec[f"{k}_edge_index"], list) or spec[f"{k}_edge_index"]):
                spec[f"{k}_edge_index"], spec[f"{k}_features"] = _apply_csr_cache(graph_dir, k, spec[f"{k}_edge_index"],
                                                                                  spec[f"{k}_features"])
    return spec


def _graph_spec_version(graph_dir):
    """``spec_version`` of the directory's metainfo file, ``"legacy"`` if the file is missing (reference
    utils/graph.py:227-262); a file without the entry, or an unknown version, is an error (:256-274)."""
    path = os.path.join(graph_dir, METAINFO_FILENAME)
    if not os.path.isfile(path):
        return "legacy"
    import yaml

    with open(path, encoding="utf-8") as f:
        meta = yaml.safe_load(f)
    version = None if meta is None else meta.get("spec_version")
    if version is None:
        raise ValueError(f"{METAINFO_FILENAME} is missing 'spec_version' entry")
    if str(version) != GRAPH_SPEC_VERSION:
        raise ValueError(f"Unsupported graph spec version {version!r} in {METAINFO_FILENAME}")
    return str(version)


def _zero_index(ei):
    """reference utils/graph.py:21-35"""
    return ei - ei.min(dim=1, keepdim=True)[0]


def _load_graph_raw(graph_dir, grid_xy):
    def ld(fn):
        return torch.load(os.path.join(graph_dir, fn), map_location="cpu", weights_only=True)

    legacy = _graph_spec_version(graph_dir) == "legacy"
    m2m_ei = list(ld("m2m_edge_index.pt"))
    m2m_f, mesh_f = ld("m2m_features.pt"), ld("mesh_features.pt")
    g2m_ei, m2g_ei = ld("g2m_edge_index.pt"), ld("m2g_edge_index.pt")
    up_ei = down_ei = None
    hier = len(m2m_ei) > 1
    if hier:
        up_ei, down_ei = list(ld("mesh_up_edge_index.pt")), list(ld("mesh_down_edge_index.pt"))
    if legacy:
        # legacy directories label the nodes of all sets with one offset numbering and hold normalised mesh
        # coordinates: zero-index every edge set on load (reference utils/graph.py:313-328, :368-374; the
        # g2m / m2g rules :38-142 — not every mesh node needs to appear in them)
        m2m_ei = [_zero_index(e) for e in m2m_ei]
        n_mesh = sum(m.shape[0] for m in mesh_f)
        mins = m2g_ei.min(dim=1)[0]
        if bool(mins[0] < mins[1]):  # mesh nodes carry the first labels
            g2m_ei = torch.stack((g2m_ei[0] - n_mesh, g2m_ei[1]))
            m2g_ei = torch.stack((m2g_ei[0], m2g_ei[1] - n_mesh))
        else:
            g2m_ei = torch.stack((g2m_ei[0], g2m_ei[1] - (g2m_ei[0].max() + 1)))
            m2g_ei = torch.stack((m2g_ei[0] - (m2g_ei[1].max() + 1), m2g_ei[1]))
        if hier:
            up_ei, down_ei = [_zero_index(e) for e in up_ei], [_zero_index(e) for e in down_ei]
    if int(m2g_ei.min()) < 0 or int(g2m_ei.min()) < 0:
        raise ValueError("negative node index in g2m / m2g edge_index")
    spec = {"hierarchical": hier, "legacy": legacy, "grid_xy": torch.as_tensor(grid_xy, dtype=torch.float32).reshape(-1, 2),
            "g2m_edge_index": g2m_ei, "m2g_edge_index": m2g_ei,
            "g2m_features": ld("g2m_features.pt"), "m2g_features": ld("m2g_features.pt")}
    if hier:
        spec.update(m2m_edge_index=m2m_ei, m2m_features=list(m2m_f), mesh_static_features=list(mesh_f),
                    mesh_up_edge_index=up_ei, mesh_up_features=list(ld("mesh_up_features.pt")),
                    mesh_down_edge_index=down_ei,
                    mesh_down_features=list(ld("mesh_down_features.pt")))
    else:
        spec.update(m2m_edge_index=m2m_ei[0], m2m_features=m2m_f[0], mesh_static_features=mesh_f[0],
                    mesh_up_edge_index=[], mesh_up_features=[], mesh_down_edge_index=[], mesh_down_features=[])
    return spec


_EDGE_SETS = ("g2m", "m2g", "m2m", "mesh_up", "mesh_down")


def save_graph_csr_cache(graph_dir):
    """Cached CSR next to the reference's ``.pt`` files: for every edge set of the graph
    directory write ``<set>_csr.pt`` = {"order": stable receiver-sort permutation of the stored edge order (int64),
    "rowptr": receiver CSR offsets (int32), "n_edges": E} — per level for the list-valued sets.  ``load_graph``
    applies it, so that edges and their static features arrive in the kernels' storage order (receiver-sorted CSR)
    without sorting at model construction."""
    def ld(fn):
        return torch.load(os.path.join(graph_dir, fn), map_location="cpu", weights_only=True)

    for k in _EDGE_SETS:
        fn = os.path.join(graph_dir, f"{k}_edge_index.pt")
        if not os.path.isfile(fn):
            continue
        ei = ld(f"{k}_edge_index.pt")
        as_list = isinstance(ei, (list, tuple))
        entries = []
        for e in (ei if as_list else [ei]):
            rcv = e[1].long()
            order = torch.sort(rcv, stable=True).indices
            n_rec = int(rcv.max()) + 1 if rcv.numel() else 0
            rowptr = torch.zeros(n_rec + 1, dtype=torch.int32)
            rowptr[1:] = torch.cumsum(torch.bincount(rcv, minlength=n_rec), 0).to(torch.int32)
            entries.append({"order": order, "rowptr": rowptr, "n_edges": int(e.shape[1])})
        torch.save(entries if as_list else entries[0], os.path.join(graph_dir, f"{k}_csr.pt"))


def _apply_csr_cache(graph_dir, name, edge_index, features):
    """Permute one edge set (or list of levels) into its cached receiver-sorted order; stale or missing caches
    (edge count mismatch) are ignored."""
    fn = os.path.join(graph_dir, f"{name}_csr.pt")
    if not os.path.isfile(fn):
        return edge_index, features
    cache = torch.load(fn, map_location="cpu", weights_only=True)
    as_list = isinstance(edge_index, (list, tuple))
    eis, fs = (list(edge_index), list(features)) if as_list else ([edge_index], [features])
    caches = list(cache) if isinstance(cache, (list, tuple)) else [cache]
    if len(caches) != len(eis) or any(c["n_edges"] != e.shape[1] for c, e in zip(caches, eis)):
        return edge_index, features
    eis = [e[:, c["order"]] for c, e in zip(caches, eis)]
    fs = [f[c["order"]] for c, f in zip(caches, fs)]
    return (eis, fs) if as_list else (eis[0], fs[0])


class SyntheticDatastore:
    """The handful of quantities the graph step predictors read from a reference
    ``BaseDatastore`` (models/step_predictors/base.py:60-106, graph/base.py:86-135), filled
    with synthetic values of the right shape: static grid features
    ~N(0,1) (seed 123), one-step difference statistics mean 0 / std 1, rectangular boundary
    frame."""

    def __init__(self, spec, d_state=17, d_forcing=18, d_static=4, boundary_width=10, seed=123):
        Nx, Ny = spec["grid_shape"]
        self.grid_shape = (Nx, Ny)
        self.num_grid_nodes = Nx * Ny
        self.num_state_vars = d_state
        self.num_forcing_vars = d_forcing  # already includes the past/future window
        self.num_static_vars = d_static
        g = torch.Generator().manual_seed(seed)
        self.grid_static_features = torch.randn(self.num_grid_nodes, d_static, generator=g)
        self.state_diff_mean = torch.zeros(d_state)
        self.state_diff_std = torch.ones(d_state)
        # standardisation statistics and names of the state variables (used by output clamping only)
        self.state_mean = torch.zeros(d_state)
        self.state_std = torch.ones(d_state)
        self.state_var_names = [f"var{i}" for i in range(d_state)]
        m = torch.zeros(Nx, Ny)
        w = min(boundary_width, max(1, min(Nx, Ny) // 4))
        m[:w, :] = 1
        m[-w:, :] = 1
        m[:, :w] = 1
        m[:, -w:] = 1
        self.boundary_mask = m.reshape(-1, 1)
        gxy = spec["grid_xy"]
        self.grid_xy = gxy
        self.grid_xy_max_span = float(max(gxy[:, 0].max() - gxy[:, 0].min(), gxy[:, 1].max() - gxy[:, 1].min()))

    @property
    def grid_input_dim(self):
        # reference utils/graph.py:507-512: 2*state + static + forcing window
        return 2 * self.num_state_vars + self.num_static_vars + self.num_forcing_vars

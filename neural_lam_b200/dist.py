"""Node-partitioned rollout across the GPUs of one box.

The reference only offers data-parallel replicas (Lightning DDP, reference README.md:486-514);
partitioning the graph is new here.  Every receiver's update depends only on its own row, its
incoming edges and the sender rows of those edges, so the path shards with ONE exchange step per
InteractionNet call:

  * node sets (grid nodes, mesh nodes) are split into contiguous index ranges — spatial strips,
    because grid node (i,j) -> i*Ny+j and mesh node (a,b) -> a*n+b;
  * a rank owns the receivers of its range and every edge whose RECEIVER it owns (edge rows never
    move) and keeps a halo copy of the off-rank SENDER rows those edges read;
  * before each GNN call the owners pack the boundary sender rows (``nlam_gather_rows``) and the
    ranks exchange them with grouped point-to-point sends/receives (``batch_isend_irecv``: one
    ncclGroup of ncclSend/ncclRecv over NVLink — an all-to-all-v; multiscale / hierarchical long
    edges make the pattern non-nearest-neighbour), then run the unchanged kernels on
    ``[owned rows | halo rows]``.

One process per GPU (``torch.distributed``, backend nccl; gloo for the CPU tests of the plan /
exchange logic).

Two exchange transports:
  * ``SymmHaloExchanger`` (CUDA, default): every rank keeps its extended sender buffers ``[own rows | halo rows]`` in
    SYMMETRIC MEMORY (``torch.distributed._symmetric_memory``: the peers' buffers are mapped into this process over
    NVLink / NVSwitch).  One launch of ``halo_push_kernel`` (libnlam_b200) copies the own rows to the front of the local
    buffer and stores the boundary rows straight into the peers' halo regions; one device-side barrier later the edge
    kernels read ``[own | halo]`` in place — no pack buffer, no NCCL send/recv, no ``torch.cat``, and the whole forecast
    step (pushes and barriers included) is captured in the CUDA graph.  Buffers are double-buffered per edge set, so
    one barrier per exchange suffices.
  * ``P2PExchanger`` (gloo / fallback): ``HaloPlan.exchange`` — pack, grouped isend/irecv, concatenate.
``partition_model`` builds ANY of the graph models (GraphLAM, HiLAM) on the rank's local sub-graph — every node set
(grid, every mesh level) is split into contiguous ranges, every edge set gets a ``HaloPlan`` — and attaches the exchange
to each InteractionNet, so the model code itself runs unchanged on ``[own rows]`` tensors.
"""
import ctypes

import torch
import torch.distributed as dist

from . import models, ops
from .synthetic import normalize_graph


def split_bounds(n, world):
    """Contiguous, near-equal index ranges: bounds[r] .. bounds[r+1]."""
    return [(n * r) // world for r in range(world + 1)]


class HaloPlan:
    """Static exchange plan of ONE edge set for ONE rank.

    edge_index : (2, E) global, zero-based (senders / receivers in their own node sets)
    send_bounds / recv_bounds : node-range ownership of the sender / receiver node sets
    Local numbering: receivers -> 0..n_rec_own-1; senders -> ``[owned rows | halo rows grouped by
    owner rank, ascending global id]``.  Every rank can build every rank's plan from the global
    graph, so setting up needs no communication.
    """

    def __init__(self, edge_index, send_bounds, recv_bounds, rank, world):
        self.rank, self.world = rank, world
        ei = edge_index.cpu().long()
        snd, rcv = ei[0], ei[1]
        r0, r1 = recv_bounds[rank], recv_bounds[rank + 1]
        s0, s1 = send_bounds[rank], send_bounds[rank + 1]
        self.n_rec_own, self.n_send_own = r1 - r0, s1 - s0
        mask = (rcv >= r0) & (rcv < r1)
        self.edge_ids = torch.nonzero(mask, as_tuple=False).squeeze(1)  # global ids of the local edges
        lsnd, lrcv = snd[mask], rcv[mask] - r0
        sb = torch.tensor(send_bounds)
        owner = torch.bucketize(lsnd, sb[1:], right=True)
        # halo rows I need, grouped by owner
        self.recv_ids = []  # per peer: ascending global sender ids received from that peer
        for peer in range(world):
            if peer == rank:
                self.recv_ids.append(torch.empty(0, dtype=torch.long))
            else:
                self.recv_ids.append(torch.unique(lsnd[owner == peer]))
        self.n_halo = int(sum(t.numel() for t in self.recv_ids))
        # local sender index of every local edge
        local = torch.empty_like(lsnd)
        own = owner == rank
        local[own] = lsnd[own] - s0
        off = self.n_send_own
        for peer in range(world):
            ids = self.recv_ids[peer]
            if ids.numel():
                sel = owner == peer
                local[sel] = off + torch.searchsorted(ids, lsnd[sel])
                off += ids.numel()
        self.local_edge_index = torch.stack([local, lrcv])
        # rows each peer needs from me (their recv_ids[rank]), as local indices into my owned range
        self.send_ids = []
        for peer in range(world):
            if peer == rank:
                self.send_ids.append(torch.empty(0, dtype=torch.long))
                continue
            p0, p1 = recv_bounds[peer], recv_bounds[peer + 1]
            pm = (rcv >= p0) & (rcv < p1)
            need = torch.unique(snd[pm])
            mine = need[(need >= s0) & (need < s1)]
            self.send_ids.append(mine - s0)
        self._dev_send = None

    # ------------------------------------------------------------------ exchange
    def to(self, device):
        self._dev_send = [t.to(device=device, dtype=torch.int32).contiguous() for t in self.send_ids]
        self._dev_send64 = [t.to(device) for t in self.send_ids]
        return self

    def exchange(self, x_own, group=None):
        """(B, n_send_own, H) owned sender rows -> (B, n_send_own + n_halo, H) extended rows."""
        if self.world == 1 or (self.n_halo == 0 and all(t.numel() == 0 for t in self.send_ids)):
            return x_own
        if self._dev_send is None or (self._dev_send64[0].device != x_own.device):
            self.to(x_own.device)
        B, _, H = x_own.shape
        x_own = x_own.contiguous()
        sends, recvs, opsl = [], [], []
        for peer in range(self.world):
            if peer == self.rank:
                continue
            if self.send_ids[peer].numel():
                if x_own.is_cuda:
                    buf = ops.gather_rows(x_own, self._dev_send[peer])  # pack kernel
                else:
                    buf = x_own.index_select(1, self._dev_send64[peer])
                sends.append(buf)
                opsl.append(dist.P2POp(dist.isend, buf, peer, group=group))
            n = self.recv_ids[peer].numel()
            if n:
                rb = torch.empty(B, n, H, dtype=x_own.dtype, device=x_own.device)
                recvs.append(rb)
                opsl.append(dist.P2POp(dist.irecv, rb, peer, group=group))
        if opsl:
            for req in dist.batch_isend_irecv(opsl):
                req.wait()
        return torch.cat([x_own] + recvs, dim=1) if recvs else x_own

    def halo_bytes(self, B, H):
        """(bytes sent, bytes received) per exchange for batch B, width H (fp32)."""
        ns = sum(t.numel() for t in self.send_ids)
        return 4 * B * H * ns, 4 * B * H * self.n_halo


class LocalDatastore:
    """The rank's slice of a datastore (grid rows g0:g1): what the step predictors and the ARForecaster read."""

    def __init__(self, ds, g0, g1):
        self.grid_static_features = ds.grid_static_features[g0:g1]
        self.boundary_mask = ds.boundary_mask[g0:g1]
        self.num_grid_nodes = g1 - g0
        for k in ("num_state_vars", "num_forcing_vars", "num_static_vars", "state_diff_mean", "state_diff_std", "state_mean",
                  "state_std", "state_var_names", "grid_input_dim"):
            setattr(self, k, getattr(ds, k))


def _edge_sets(g):
    """[(name, edge_index, sender node set, receiver node set)] of a (normalised) graph spec; node sets are named
    "grid", "mesh0", "mesh1", ..."""
    sets = [("g2m", g["g2m_edge_index"], "grid", "mesh0"), ("m2g", g["m2g_edge_index"], "mesh0", "grid")]
    if g["hierarchical"]:
        for l, ei in enumerate(g["m2m_edge_index"]):
            sets.append((f"m2m{l}", ei, f"mesh{l}", f"mesh{l}"))
        for l, ei in enumerate(g["mesh_up_edge_index"]):
            sets.append((f"up{l}", ei, f"mesh{l}", f"mesh{l + 1}"))
        for l, ei in enumerate(g["mesh_down_edge_index"]):
            sets.append((f"down{l}", ei, f"mesh{l + 1}", f"mesh{l}"))
    else:
        sets.append(("m2m0", g["m2m_edge_index"], "mesh0", "mesh0"))
    return sets


def partition_graph(graph, datastore, rank, world):
    """(local graph spec, LocalDatastore, {edge set: [HaloPlan of every rank]}, {node set: bounds}) — every rank can
    build every rank's plan from the global graph, so setting up needs no communication."""
    g = graph if graph.get("normalized") else normalize_graph(graph)
    hier = g["hierarchical"]
    sizes = {"grid": datastore.grid_static_features.shape[0]}
    mesh = g["mesh_static_features"] if hier else [g["mesh_static_features"]]
    for l, m in enumerate(mesh):
        sizes[f"mesh{l}"] = m.shape[0]
    bounds = {k: split_bounds(n, world) for k, n in sizes.items()}
    plans = {}
    for name, ei, ss, rs in _edge_sets(g):
        plans[name] = [HaloPlan(ei, bounds[ss], bounds[rs], r, world) for r in range(world)]
        mine = plans[name][rank]
        assert int(mine.local_edge_index[1].max()) + 1 == mine.n_rec_own, f"{name}: trailing receiver without edges"
    local = dict(g)
    local["normalized"] = True

    def cut(name, key_ei, key_f, l=None):
        plan = plans[name][rank]
        f = g[key_f] if l is None else g[key_f][l]
        return plan.local_edge_index, f[plan.edge_ids]

    local["g2m_edge_index"], local["g2m_features"] = cut("g2m", "g2m_edge_index", "g2m_features")
    local["m2g_edge_index"], local["m2g_features"] = cut("m2g", "m2g_edge_index", "m2g_features")
    if hier:
        L = len(mesh)
        for key, pre, n in (("m2m", "m2m", L), ("mesh_up", "up", L - 1), ("mesh_down", "down", L - 1)):
            pairs = [cut(f"{pre}{l}", f"{key}_edge_index", f"{key}_features", l) for l in range(n)]
            local[f"{key}_edge_index"] = [p[0] for p in pairs]
            local[f"{key}_features"] = [p[1] for p in pairs]
        local["mesh_static_features"] = [m[bounds[f"mesh{l}"][rank]:bounds[f"mesh{l}"][rank + 1]] for l, m in enumerate(mesh)]
    else:
        local["m2m_edge_index"], local["m2m_features"] = cut("m2m0", "m2m_edge_index", "m2m_features")
        local["mesh_static_features"] = mesh[0][bounds["mesh0"][rank]:bounds["mesh0"][rank + 1]]
    lds = LocalDatastore(datastore, bounds["grid"][rank], bounds["grid"][rank + 1])
    return local, lds, plans, bounds


class P2PExchanger:
    """Pack + grouped isend/irecv + concatenate (``HaloPlan.exchange``): the gloo / fallback transport."""

    def __init__(self, plans, rank, world, group=None):
        self.plans, self.rank, self.world, self.group = plans, rank, world, group

    def exchange(self, name, x_own):
        return self.plans[name][self.rank].exchange(x_own, self.group)


class SymmHaloExchanger:
    """Extended sender buffers in symmetric memory + ``nlam_halo_push`` + one device barrier per exchange."""

    def __init__(self, plans, rank, world, group=None):
        self.plans, self.rank, self.world = plans, rank, world
        self.group = group if group is not None else dist.group.WORLD
        self._arena = None
        self._key = None
        self._count = {}

    def _setup(self, B, H, device):
        import torch.distributed._symmetric_memory as symm_mem

        names = sorted(self.plans)
        rows_max = {n: max(pl.n_send_own + pl.n_halo for pl in self.plans[n]) for n in names}
        self._rows = rows_max
        off, self._off = 0, {}
        for n in names:
            for k in (0, 1):
                self._off[(n, k)] = off
                off += B * rows_max[n] * H
        arena = symm_mem.empty(off, dtype=torch.float32, device=device)
        self._hdl = symm_mem.rendezvous(arena, self.group)
        self._arena = arena
        base = [int(p) for p in self._hdl.buffer_ptrs]
        self._tab = {}
        for n in names:
            me = self.plans[n][self.rank]
            send_rows = torch.cat([me.send_ids[p] for p in range(self.world)]).to(torch.int32)
            ptr = [0]
            for p in range(self.world):
                ptr.append(ptr[-1] + me.send_ids[p].numel())
            dst_off = []
            for p in range(self.world):
                pl = self.plans[n][p]  # where my rows start in peer p's halo region
                dst_off.append(pl.n_send_own + int(sum(pl.recv_ids[q].numel() for q in range(self.rank))))
            entry = {"send_rows": send_rows.to(device), "send_ptr": torch.tensor(ptr, dtype=torch.int32, device=device),
                     "dst_off": torch.tensor(dst_off, dtype=torch.int32, device=device), "n_send": ptr[-1],
                     "n_own": me.n_send_own, "n_ext": me.n_send_own + me.n_halo}
            for k in (0, 1):
                entry[("peers", k)] = torch.tensor([b + 4 * self._off[(n, k)] for b in base], dtype=torch.int64, device=device)
            self._tab[n] = entry
        self._key = (B, H, device)
        self._hdl.barrier(channel=0)

    def exchange(self, name, x_own):
        from . import _lib

        B, n_own, H = x_own.shape
        if self._key != (B, H, x_own.device):
            self._setup(B, H, x_own.device)
        t = self._tab[name]
        assert n_own == t["n_own"], (name, n_own, t["n_own"])
        k = self._count.get(name, 0) & 1
        self._count[name] = self._count.get(name, 0) + 1
        rows = self._rows[name]
        o = self._off[(name, k)]
        ext = self._arena[o:o + B * rows * H].view(B, rows, H)
        x = x_own if x_own.stride(-1) == 1 and x_own.stride(1) == H else x_own.contiguous()
        xbs = x.stride(0) if B > 1 else 0
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().nlam_halo_push(
                x.data_ptr(), xbs, n_own, ext.data_ptr(), rows * H, t[("peers", k)].data_ptr(), t["send_rows"].data_ptr(),
                t["send_ptr"].data_ptr(), t["dst_off"].data_ptr(), t["n_send"], self.world, B, H,
                ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        self._hdl.barrier(channel=0)   # every peer's rows have landed in my halo region (device-side, capturable)
        return ext[:, :t["n_ext"]]

    def halo_bytes(self, name, B, H):
        pl = self.plans[name][self.rank]
        return pl.halo_bytes(B, H)


def _gnn_edge_sets(model):
    """[(InteractionNet module, edge-set name)] of a GraphLAM / HiLAM model."""
    out = [(model.g2m_gnn, "g2m"), (model.m2g_gnn, "m2g")]
    if not model.hierarchical:
        out += [(layer, "m2m0") for layer in model.processor.children()]
        return out
    for l, gnn in enumerate(model.mesh_init_gnns):
        out.append((gnn, f"up{l}"))
    for l, gnn in enumerate(model.mesh_read_gnns):
        out.append((gnn, f"down{l}"))
    for name, pre in (("mesh_down_gnns", "down"), ("mesh_up_gnns", "up"), ("mesh_down_same_gnns", "m2m"), ("mesh_up_same_gnns", "m2m")):
        for per_layer in getattr(model, name):
            for l, gnn in enumerate(per_layer):
                out.append((gnn, f"{pre}{l}"))
    return out


def partition_model(model_cls, datastore, graph, rank, world, group=None, transport="auto", **kwargs):
    """Build ``model_cls`` (``models.GraphLAM`` / ``models.HiLAM``) on this rank's part of the graph.  Inputs / outputs
    of the returned model are the rank's OWN grid rows ``(B, G_own, d)``; parameters are replicated (same seed ->
    identical weights on every rank).  Extra attributes: ``plans``, ``bounds``, ``local_datastore`` (for the
    ARForecaster), ``exchanger``, ``own_grid_slice()``, ``halo_bytes_per_step(B)``."""
    if kwargs.get("output_std") or kwargs.get("output_clamping_lower") or kwargs.get("output_clamping_upper"):
        raise NotImplementedError("partition_model: output_std / output clamping are not supported on the partitioned path")
    local, lds, plans, bounds = partition_graph(graph, datastore, rank, world)
    model = model_cls(lds, local, **kwargs)
    if transport == "auto":
        transport = "symm" if (torch.cuda.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl") else "p2p"
    ex = (SymmHaloExchanger if transport == "symm" else P2PExchanger)(plans, rank, world, group)
    for gnn, name in _gnn_edge_sets(model):
        gnn._halo = (lambda x, _n=name: ex.exchange(_n, x))
    model.rank, model.world, model.group = rank, world, group
    model.plans, model.bounds, model.local_datastore, model.exchanger = plans, bounds, lds, ex
    g0, g1 = bounds["grid"][rank], bounds["grid"][rank + 1]
    model.own_grid_slice = lambda: slice(g0, g1)

    def halo_bytes_per_step(B):
        H = model.hidden_dim
        tot = [0, 0]
        for _, name in _gnn_edge_sets(model):
            s_, r_ = plans[name][rank].halo_bytes(B, H)
            tot[0] += s_
            tot[1] += r_
        return tuple(tot)

    model.halo_bytes_per_step = halo_bytes_per_step
    return model


def PartitionedGraphLAM(datastore, graph, rank, world, group=None, **kwargs):
    """GraphLAM on this rank's strip of the graph (see ``partition_model``)."""
    return partition_model(models.GraphLAM, datastore, graph, rank, world, group=group, **kwargs)


def PartitionedHiLAM(datastore, graph, rank, world, group=None, **kwargs):
    """HiLAM with EVERY mesh level split into strips (see ``partition_model``)."""
    return partition_model(models.HiLAM, datastore, graph, rank, world, group=group, **kwargs)

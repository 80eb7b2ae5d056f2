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
"""
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


class PartitionedGraphLAM(models.GraphLAM):
    """GraphLAM whose grid and mesh nodes are split into ``world`` contiguous strips; this rank
    holds the rows it owns, the edges whose receivers it owns, and exchanges boundary sender rows
    before each of the 1 + P + 1 InteractionNet calls of a step.  Parameters are replicated (same
    seed -> identical weights on every rank); inputs / outputs are the rank's OWN grid rows."""

    def __init__(self, datastore, graph, rank, world, group=None, **kwargs):
        g = graph if graph.get("normalized") else normalize_graph(graph)
        assert not g["hierarchical"]
        G = datastore.grid_static_features.shape[0]
        M = g["mesh_static_features"].shape[0]
        grid_bounds = split_bounds(G, world)
        mesh_bounds = split_bounds(M, world)
        plans = {
            "g2m": HaloPlan(g["g2m_edge_index"], grid_bounds, mesh_bounds, rank, world),
            "m2m": HaloPlan(g["m2m_edge_index"], mesh_bounds, mesh_bounds, rank, world),
            "m2g": HaloPlan(g["m2g_edge_index"], mesh_bounds, grid_bounds, rank, world),
        }
        for name, plan in plans.items():
            assert int(plan.local_edge_index[1].max()) + 1 == plan.n_rec_own, f"{name}: trailing receiver without edges"
        g0, g1 = grid_bounds[rank], grid_bounds[rank + 1]
        m0, m1 = mesh_bounds[rank], mesh_bounds[rank + 1]
        local = dict(g)
        for name, plan in plans.items():
            local[f"{name}_edge_index"] = plan.local_edge_index
            local[f"{name}_features"] = g[f"{name}_features"][plan.edge_ids]
        local["mesh_static_features"] = g["mesh_static_features"][m0:m1]
        local["normalized"] = True

        class _LocalStore:
            pass

        ds = _LocalStore()
        ds.grid_static_features = datastore.grid_static_features[g0:g1]
        ds.num_state_vars = datastore.num_state_vars
        ds.state_diff_mean, ds.state_diff_std = datastore.state_diff_mean, datastore.state_diff_std
        ds.grid_input_dim = datastore.grid_input_dim
        ds.state_var_names, ds.state_mean, ds.state_std = datastore.state_var_names, datastore.state_mean, datastore.state_std
        super().__init__(ds, local, **kwargs)
        if self.output_std or self.clamps_output:
            raise NotImplementedError("PartitionedGraphLAM: output_std / output clamping are not supported on the "
                                      "partitioned inference path")
        self.rank, self.world, self.group = rank, world, group
        self.grid_bounds, self.mesh_bounds = grid_bounds, mesh_bounds
        self.plans = plans
        self.boundary_mask_local = datastore.boundary_mask[g0:g1].float()

    def own_grid_slice(self):
        return slice(self.grid_bounds[self.rank], self.grid_bounds[self.rank + 1])

    @torch.no_grad()
    def forward(self, prev_state, prev_prev_state, forcing):
        """Inputs / output: this rank's own grid rows ``(B, G_own, d)``."""
        B = prev_state.shape[0]
        ex = self.expand_to_batch
        grid_emb = self.grid_embedder.apply_rows([prev_state, prev_prev_state, forcing, ex(self.grid_static_features, B)])
        st = self.static_embeddings()
        grid_ext = self.plans["g2m"].exchange(grid_emb, self.group)          # halo of grid sender rows
        mesh_rep = self.g2m_gnn(grid_ext, ex(st["mesh_emb"], B), ex(st["g2m_emb"], B))
        grid_rep = self.encoding_grid_mlp.apply_rows([grid_emb], res=grid_emb)
        edge = ex(st["m2m_emb"], B)
        for layer in self.processor.children():
            mesh_ext = self.plans["m2m"].exchange(mesh_rep, self.group)      # halo of mesh sender rows
            mesh_rep, edge = layer(mesh_ext, mesh_rep, edge)
        mesh_ext = self.plans["m2g"].exchange(mesh_rep, self.group)
        grid_rep = self.m2g_gnn(mesh_ext, grid_rep, ex(st["m2g_emb"], B))
        net_output = self.output_map(grid_rep)
        return ops.step_epilogue(net_output, prev_state, None, None, self.diff_std, self.diff_mean), None

    def halo_bytes_per_step(self, B):
        H = self.hidden_dim
        P = len(self.processor)
        tot = [0, 0]
        for name, mult in (("g2m", 1), ("m2m", P), ("m2g", 1)):
            s, r = self.plans[name].halo_bytes(B, H)
            tot[0] += mult * s
            tot[1] += mult * r
        return tuple(tot)

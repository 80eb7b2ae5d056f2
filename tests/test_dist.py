"""Node-partitioned path: halo plans + exchange on 2 ranks.

CPU (gloo, world_size 2): every rank builds its HaloPlan from the global graph, exchanges the
boundary sender rows with batched isend/irecv and evaluates its receivers with the CPU oracle; the
gathered result must equal the oracle on the whole graph.  GPU (nccl, 2 GPUs; skipped on a 1-GPU
box): PartitionedGraphLAM on the kernels vs the single-GPU GraphLAM.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_lam_b200 import dist as nd
from neural_lam_b200 import models, synthetic
from oracle import reference_port as rp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _graph(ns, nr, ne, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, ns, (ne,), generator=g), torch.randint(0, nr, (ne,), generator=g)])
    ei[1, :nr] = torch.arange(nr)  # every receiver has an edge
    return ei[:, torch.sort(ei[1], stable=True).indices]


def _cpu_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        H, B = 8, 2
        results = {}
        for name, ns, nr, ne, same in (("m2m", 60, 60, 500, True), ("g2m", 300, 37, 700, False)):
            ei = _graph(ns, nr, ne, 3)
            torch.manual_seed(1)
            params = {}
            for pre, kin in (("edge_mlp", 3 * H), ("aggr_mlp", 2 * H)):
                params[f"{pre}.0.weight"] = torch.randn(H, kin) * 0.2
                params[f"{pre}.0.bias"] = torch.randn(H) * 0.1
                params[f"{pre}.2.weight"] = torch.randn(H, H) * 0.2
                params[f"{pre}.2.bias"] = torch.randn(H) * 0.1
                params[f"{pre}.3.weight"] = 1 + 0.1 * torch.randn(H)
                params[f"{pre}.3.bias"] = 0.1 * torch.randn(H)
            send = torch.randn(B, ns, H)
            rec = send if same else torch.randn(B, nr, H)
            edge = torch.randn(B, ne, H)
            want_rec, want_edge = rp.interaction_net(params, ei, send, rec, edge)
            sb, rb = nd.split_bounds(ns, world), nd.split_bounds(nr, world)
            plan = nd.HaloPlan(ei, sb, rb, rank, world)
            x_own = send[:, sb[rank]:sb[rank + 1]].contiguous()
            x_ext = plan.exchange(x_own)
            assert x_ext.shape[1] == plan.n_send_own + plan.n_halo
            # the halo rows are exactly the global rows they stand for
            off = plan.n_send_own
            for peer in range(world):
                ids = plan.recv_ids[peer]
                if ids.numel():
                    assert torch.equal(x_ext[:, off:off + ids.numel()], send[:, ids])
                    off += ids.numel()
            got_rec, got_edge = rp.interaction_net(params, plan.local_edge_index, x_ext,
                                                   rec[:, rb[rank]:rb[rank + 1]], edge[:, plan.edge_ids])
            torch.testing.assert_close(got_rec, want_rec[:, rb[rank]:rb[rank + 1]], rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(got_edge, want_edge[:, plan.edge_ids], rtol=1e-5, atol=1e-5)
            gathered = [torch.empty(B, rb[r + 1] - rb[r], H) for r in range(world)]
            dist.all_gather(gathered, got_rec.contiguous()) if len({t.shape for t in gathered}) == 1 else None
            results[name] = (plan.n_halo, plan.halo_bytes(B, H))
        q.put((rank, "ok", results))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_halo_plan_and_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in out:
        assert status == "ok", f"rank {rank}: {info}"
    # both ranks exchange something on these random graphs
    assert all(info["m2m"][0] > 0 for _, _, info in out)


def test_halo_plan_covers_every_edge_once():
    ei = _graph(200, 50, 900, 7)
    world = 4
    sb, rb = nd.split_bounds(200, world), nd.split_bounds(50, world)
    seen = torch.zeros(900, dtype=torch.int32)
    for r in range(world):
        plan = nd.HaloPlan(ei, sb, rb, r, world)
        seen[plan.edge_ids] += 1
        # local senders resolve to the right global ids
        glob = torch.cat([torch.arange(sb[r], sb[r + 1])] + [plan.recv_ids[p] for p in range(world)])
        assert torch.equal(glob[plan.local_edge_index[0]], ei[0, plan.edge_ids])
        assert torch.equal(plan.local_edge_index[1] + rb[r], ei[1, plan.edge_ids])
        # what peers send me is what I expect to receive
        for peer in range(world):
            if peer != r:
                pp = nd.HaloPlan(ei, sb, rb, peer, world)
                assert torch.equal(pp.send_ids[r] + sb[peer], plan.recv_ids[peer])
    assert torch.all(seen == 1)


def test_partition_graph_hierarchical_local_graphs_are_consistent():
    """``partition_graph`` on a 3-level hierarchical graph: for every edge set and every rank the local edge_index
    resolves to the global edges the rank owns, the local features are the owned edges' features, every edge is
    owned exactly once, and the symmetric-exchange bookkeeping (where my rows land in a peer's halo region) agrees
    with what the peer expects."""
    spec = synthetic.make_graph_spec(81, 81, hierarchical=True)
    ds = synthetic.SyntheticDatastore(spec, d_state=5, d_forcing=6, d_static=1, boundary_width=2)
    g = synthetic.normalize_graph(spec)
    world = 4
    parts = [nd.partition_graph(spec, ds, r, world) for r in range(world)]
    sets = {name: (ei, ss, rs) for name, ei, ss, rs in nd._edge_sets(g)}
    assert set(sets) == {"g2m", "m2g", "m2m0", "m2m1", "m2m2", "up0", "up1", "down0", "down1"}
    for name, (ei, ss, rs) in sets.items():
        seen = torch.zeros(ei.shape[1], dtype=torch.int32)
        for r in range(world):
            local, lds, plans, bounds = parts[r]
            plan = plans[name][r]
            sb, rb = bounds[ss], bounds[rs]
            seen[plan.edge_ids] += 1
            glob = torch.cat([torch.arange(sb[r], sb[r + 1])] + [plan.recv_ids[p] for p in range(world)])
            assert torch.equal(glob[plan.local_edge_index[0]], ei[0, plan.edge_ids])
            assert torch.equal(plan.local_edge_index[1] + rb[r], ei[1, plan.edge_ids])
            # my rows land in peer p's halo region right after the rows of the lower ranks
            for peer in range(world):
                if peer == r or plan.send_ids[peer].numel() == 0:
                    continue
                pl = plans[name][peer]
                off = pl.n_send_own + sum(pl.recv_ids[q].numel() for q in range(r))
                ext_ids = torch.cat([torch.arange(sb[peer], sb[peer + 1])] + [pl.recv_ids[q] for q in range(world)])
                n = plan.send_ids[peer].numel()
                assert torch.equal(ext_ids[off:off + n], plan.send_ids[peer] + sb[r])
        assert torch.all(seen == 1), name
    local, lds, plans, bounds = parts[1]
    assert lds.grid_static_features.shape[0] == bounds["grid"][2] - bounds["grid"][1]
    torch.testing.assert_close(local["mesh_up_features"][0], g["mesh_up_features"][0][plans["up0"][1].edge_ids])
    assert [m.shape[0] for m in local["mesh_static_features"]] == [bounds[f"mesh{l}"][2] - bounds[f"mesh{l}"][1] for l in range(3)]


def _gpu_worker(rank, world, port, q, kind):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = torch.device("cuda", rank)
        hier = kind == "hi_lam"
        spec = synthetic.make_graph_spec(81, 81, hierarchical=hier) if hier else synthetic.make_graph_spec(60, 54)
        ds = synthetic.SyntheticDatastore(spec, d_state=17, d_forcing=18, d_static=4, boundary_width=2)
        cls, pcls = (models.HiLAM, nd.PartitionedHiLAM) if hier else (models.GraphLAM, nd.PartitionedGraphLAM)
        torch.manual_seed(42)
        ref = cls(ds, spec, hidden_dim=64, processor_layers=2, math="auto").to(dev)
        torch.manual_seed(42)
        part = pcls(ds, spec, rank, world, hidden_dim=64, processor_layers=2, math="auto").to(dev)
        assert type(part.exchanger).__name__ == "SymmHaloExchanger"
        G = ref.num_grid_nodes
        gen = torch.Generator().manual_seed(5)
        B, T = 2, 3
        init = torch.randn(B, 2, G, 17, generator=gen).to(dev)
        forc = torch.randn(B, T, G, 18, generator=gen).to(dev)
        bnd = torch.randn(B, T, G, 17, generator=gen).to(dev)
        sl = part.own_grid_slice()
        fc_ref = models.ARForecaster(ref, ds).to(dev).eval()
        fc = models.ARForecaster(part, part.local_datastore).to(dev).eval()
        with torch.no_grad():
            want, _ = fc_ref(init, forc, bnd)
            own = [t[:, :, sl].contiguous() for t in (init, forc, bnd)]
            got, _ = fc(*own)                                    # eager: pushes + barriers on the stream
            graphed = fc.rollout_graphed(*own)                  # the same step replayed from CUDA graphs
        torch.cuda.synchronize()
        err = (got - want[:, :, sl]).abs().max().item()
        assert err < 2e-2, err  # same TF32 kernels, only the tile composition differs; 3 AR steps
        torch.testing.assert_close(graphed, got, rtol=1e-6, atol=1e-6)
        q.put((rank, "ok", (err, part.halo_bytes_per_step(B))))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind", ["graph_lam", "hi_lam"])
def test_partitioned_models_match_single_gpu_symmetric_memory(kind):
    """2 ranks, NCCL process group, halo rows pushed straight into the peer's symmetric-memory buffer by
    ``halo_push_kernel``: partitioned GraphLAM / HiLAM rollout (eager and CUDA-graph replay) == the single-GPU model."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in out:
        assert status == "ok", f"rank {rank}: {info}"
    print(kind, [info for _, _, info in out])

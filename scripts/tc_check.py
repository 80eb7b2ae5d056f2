"""GPU bring-up check of the tcgen05 kernels: each mode vs the fp64 oracle, with timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import ops, _lib
from oracle import reference_port as rp

dev = torch.device("cuda:0")
torch.manual_seed(0)
TF = _lib.MATH_TF32


def err(a, b):
    return (a.double().cpu() - b).abs().max().item()


def check_rowmlp(name, bp, srcs_shapes, B, res_idx=None, ln=True):
    m = nlb.make_mlp(bp, layer_norm=ln)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    srcs = [torch.randn(*s) for s in srcs_shapes]
    B_eff = max([s.shape[0] for s in srcs if s.dim() == 3], default=1)
    cat = torch.cat([s if s.dim() == 3 else s.unsqueeze(0).expand(B_eff, -1, -1) for s in srcs], dim=-1).double()
    params = {f"m.{k}": v.double() for k, v in m.state_dict().items()}
    want = rp.mlp(cat, params, "m", layer_norm=ln)
    if res_idx is not None:
        r = srcs[res_idx]
        want = want + (r if r.dim() == 3 else r.unsqueeze(0)).double()
    m = m.to(dev)
    ds = [s.to(dev) for s in srcs]
    with torch.no_grad():
        got = ops.rowmlp(m, ds, res=ds[res_idx] if res_idx is not None else None, flags=TF)
        got32 = ops.rowmlp(m, ds, res=ds[res_idx] if res_idx is not None else None, flags=_lib.MATH_FP32)
    torch.cuda.synchronize()
    if got.dim() == 2:
        got, got32 = got.unsqueeze(0), got32.unsqueeze(0)
    print(f"{name:34s} tf32 err {err(got, want):.3e}   fp32 err {err(got32, want):.3e}   shape {tuple(got.shape)}", flush=True)


def check_edge(name, ns, nr, ne, B, update_edges, aggr="sum", same=False, sorted_edges=True, expand_edge=False):
    g = torch.Generator().manual_seed(1)
    ei = torch.stack([torch.randint(0, ns, (ne,), generator=g), torch.randint(0, nr, (ne,), generator=g)])
    ei[1, -1] = nr - 1
    if sorted_edges:
        ei = ei[:, torch.sort(ei[1], stable=True).indices]
    net = nlb.InteractionNet(ei, 64, update_edges=update_edges, aggr=aggr, math="tf32")
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    send = torch.randn(B, ns, 64)
    rec = send if same else torch.randn(B, nr, 64)
    edge = torch.randn(1 if expand_edge else B, ne, 64)
    sd = {k: v.double() for k, v in net.state_dict().items()}
    out = rp.interaction_net(sd, ei, send.double(), rec.double(), edge.double().expand(B, -1, -1), aggr=aggr, update_edges=update_edges)
    net = net.to(dev)
    with torch.no_grad():
        got = net(send.to(dev), (send if same else rec).to(dev), edge.to(dev).expand(B, -1, -1))
    torch.cuda.synchronize()
    if update_edges:
        print(f"{name:34s} tf32 rec err {err(got[0], out[0]):.3e} edge err {err(got[1], out[1]):.3e}", flush=True)
    else:
        print(f"{name:34s} tf32 rec err {err(got, out):.3e}", flush=True)
    return net


print("== row mode ==", flush=True)
check_rowmlp("embed 64->64->64 LN", [64, 64, 64], [(3, 300, 64)], 3)
check_rowmlp("embed+res (encoding_grid_mlp)", [64, 64, 64], [(2, 1000, 64)], 2, res_idx=0)
check_rowmlp("node [rec|aggr] res=rec", [128, 64, 64], [(2, 777, 64), (2, 777, 64)], 2, res_idx=0)
check_rowmlp("node [rec|aggr] res=aggr", [128, 64, 64], [(2, 130, 64), (2, 130, 64)], 2, res_idx=1)
check_rowmlp("node bcast rec", [128, 64, 64], [(257, 64), (3, 257, 64)], 3, res_idx=None)
check_rowmlp("output_map 64->64->17 noLN", [64, 64, 17], [(2, 500, 64)], 2, ln=False)
check_rowmlp("grid embedder 17|17|18|4", [56, 64, 64], [(2, 400, 17), (2, 400, 17), (2, 400, 18), (400, 4)], 2)
print("== edge mode ==", flush=True)
check_edge("edge small sum upd", 50, 30, 400, 2, True)
check_edge("edge small mean noupd", 50, 30, 400, 2, False, aggr="mean")
check_edge("edge m2m-like same nodes", 200, 200, 1800, 3, True, same=True)
check_edge("edge unsorted", 60, 40, 500, 2, True, sorted_edges=False)
check_edge("edge expand()ed edge input", 60, 40, 500, 3, True, expand_edge=True)
check_edge("edge many empty receivers", 40, 600, 300, 2, True)

print("== MEPS m2m timing ==", flush=True)
from neural_lam_b200 import synthetic
spec = synthetic.make_graph_spec(268, 238)
ei = spec["m2m_edge_index"]
for B in (1, 4, 8):
    for math in ("tf32", "fp32"):
        net = nlb.InteractionNet(ei, 64, math=math).to(dev)
        mesh = torch.randn(B, 6561, 64, device=dev)
        edge = torch.randn(B, ei.shape[1], 64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        with torch.no_grad():
            for _ in range(3):
                net(mesh, mesh, edge)
            ts = []
            for _ in range(10):
                flush.zero_()
                s, e = torch.cuda.Event(True), torch.cuda.Event(True)
                s.record(); net(mesh, mesh, edge); e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
        nbytes = rp.algorithmic_bytes_inet(B, 6561, 6561, ei.shape[1], 64, True, True)
        t = sorted(ts)[len(ts) // 2]
        print(f"m2m B={B} {math}: {t*1e3:.1f} us  -> {nbytes/t/1e6:.0f} GB/s algorithmic ({nbytes/t/1e6/6566.7*100:.1f}% of measured HBM peak)", flush=True)
print("tc_check done")

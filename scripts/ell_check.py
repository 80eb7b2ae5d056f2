"""GPU check of the uniform in-degree (ELL) edge kernel vs the fp64 oracle + MEPS m2g timing.
Run twice to compare: NLAM_TC_NO_ELL=1 python scripts/ell_check.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import synthetic
from oracle import reference_port as rp

dev = torch.device("cuda:0")
torch.manual_seed(0)


def err(a, b):
    return (a.double().cpu() - b).abs().max().item()


def check(name, ns, nr, d, B, aggr="sum", expand_edge=False, bcast_rec=False):
    g = torch.Generator().manual_seed(1)
    rcv = torch.arange(nr).repeat_interleave(d)
    snd = torch.randint(0, ns, (nr * d,), generator=g)
    ei = torch.stack([snd, rcv])
    net = nlb.InteractionNet(ei, 64, update_edges=False, aggr=aggr, math="tf32")
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    send = torch.randn(B, ns, 64)
    rec = torch.randn(1 if bcast_rec else B, nr, 64)
    edge = torch.randn(1 if expand_edge else B, nr * d, 64)
    sd = {k: v.double() for k, v in net.state_dict().items()}
    want = rp.interaction_net(sd, ei, send.double(), rec.double().expand(B, -1, -1), edge.double().expand(B, -1, -1),
                              aggr=aggr, update_edges=False)
    net = net.to(dev)
    with torch.no_grad():
        got = net(send.to(dev), rec.to(dev).expand(B, -1, -1), edge.to(dev).expand(B, -1, -1))
    torch.cuda.synchronize()
    print(f"{name:40s} rec err {err(got, want):.3e}", flush=True)


check("d=4 small", 50, 30, 4, 2)
check("d=4 nr=1000 mean", 300, 1000, 4, 3, aggr="mean")
check("d=1 (mesh down)", 100, 700, 1, 2)
check("d=2 B=1", 100, 129, 2, 1)
check("d=3 expand edge", 64, 500, 3, 3, expand_edge=True)
check("d=8 bcast rec", 64, 300, 8, 2, bcast_rec=True)
check("d=5 multi-tile per CTA", 2000, 128 * 150 + 17, 5, 2)
check("d=4 window multi-tile per CTA", 120, 128 * 150 + 17, 4, 2)
check("d=1 window multi-tile", 5000, 128 * 300 + 5, 1, 3, aggr="mean")

spec = synthetic.make_graph_spec(268, 238)
ei = spec["m2g_edge_index"]
G = 268 * 238
for B in (1, 8, 32):
    net = nlb.InteractionNet(ei, 64, update_edges=False, math="tf32").to(dev)
    mesh = torch.randn(B, 6561, 64, device=dev)
    grid = torch.randn(B, G, 64, device=dev)
    edge = torch.randn(1, ei.shape[1], 64, device=dev).expand(B, -1, -1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            net(mesh, grid, edge)
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(True), torch.cuda.Event(True)
            s.record(); net(mesh, grid, edge); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    print(f"m2g B={B}: {t*1e3:.1f} us  (ELL {'off' if os.environ.get('NLAM_TC_NO_ELL') else 'on'})", flush=True)
print("ell_check done")

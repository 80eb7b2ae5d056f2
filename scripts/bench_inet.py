"""Time the launches of ONE InteractionNet call of the MEPS GraphLAM step (g2m | m2m | m2m_last | m2g), L2 flushed
between iterations, with the library's per-launch profile.  usage: python scripts/bench_inet.py m2m 32 [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import neural_lam_b200 as nlb
from neural_lam_b200 import ops, synthetic

kind = sys.argv[1] if len(sys.argv) > 1 else "m2m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
spec = synthetic.make_graph_spec(268, 238)
torch.manual_seed(0)
G, M = 268 * 238, 6561
if kind == "g2m":
    ei = spec["g2m_edge_index"]
    net = nlb.InteractionNet(ei, 64, update_edges=False).to(dev)
    send, rec = torch.randn(B, G, 64, device=dev), torch.randn(M, 64, device=dev).unsqueeze(0).expand(B, -1, -1)
    edge = torch.randn(ei.shape[1], 64, device=dev).unsqueeze(0).expand(B, -1, -1)
    call = lambda: net(send, rec, edge)
elif kind == "m2g":
    ei = spec["m2g_edge_index"]
    net = nlb.InteractionNet(ei, 64, update_edges=False).to(dev)
    send, rec = torch.randn(B, M, 64, device=dev), torch.randn(B, G, 64, device=dev)
    edge = torch.randn(ei.shape[1], 64, device=dev).unsqueeze(0).expand(B, -1, -1)
    call = lambda: net(send, rec, edge)
else:
    ei = spec["m2m_edge_index"]
    net = nlb.InteractionNet(ei, 64).to(dev)
    mesh = torch.randn(B, M, 64, device=dev)
    edge = torch.randn(B, ei.shape[1], 64, device=dev)
    if kind == "m2m":
        call = lambda: net(mesh, mesh, edge)
    elif kind == "m2m_last":
        call = lambda: net.forward_stacked(mesh, edge, first=False, last=True)
    elif kind == "m2m_inplace":
        call = lambda: net.forward_stacked(mesh, edge, first=False, last=False)
    elif kind == "m2m_first":
        e1 = edge[0].unsqueeze(0).expand(B, -1, -1)
        call = lambda: net.forward_stacked(mesh, e1, first=True, last=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(2):
        call()
    tot = {}
    for _ in range(iters):
        flush.zero_()
        with ops.profile_launches() as prof:
            call()
        for i, (n, us, nb) in enumerate(prof.rows):
            tot.setdefault((i, n), []).append((us, nb))
for (i, n), v in tot.items():
    us = sorted(x[0] for x in v)[len(v) // 2]
    print(f"{kind} B={B} launch {i} {n:32s} {us:8.1f} us  {v[0][1] / 1e6:8.1f} MB  {v[0][1] / us / 1e3:7.1f} GB/s")
print(f"{kind} B={B} total {sum(sorted(x[0] for x in v)[len(v) // 2] for v in tot.values()):.1f} us")

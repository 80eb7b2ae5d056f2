"""Time the chained grid-tail kernel (node update + output_map + step epilogue, tc9.cu) on the MEPS grid, L2 flushed;
NLAM_TC_TIMELINE=1 prints its in-kernel timeline.  usage: python scripts/bench_node_out.py [B] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import neural_lam_b200 as nlb
from neural_lam_b200 import ops
from neural_lam_b200.networks import make_mlp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
G = 268 * 238
torch.manual_seed(0)
ei = torch.stack([torch.randint(0, 100, (4 * 256,)), torch.arange(256).repeat_interleave(4)])
net = nlb.InteractionNet(ei, 64, update_edges=False).to(dev)
out_map = make_mlp([64, 64, 17], layer_norm=False).to(dev)
rec, aggr = torch.randn(B, G, 64, device=dev), torch.randn(B, G, 64, device=dev)
prev, bnd = torch.randn(B, G, 17, device=dev), torch.randn(B, G, 17, device=dev)
mask = (torch.rand(G, 1, device=dev) < 0.15).float()
std, mean = torch.ones(17, device=dev), torch.zeros(17, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty_like(prev)
with torch.no_grad():
    for _ in range(2):
        ops.node_update_step(net.aggr_mlp, out_map, rec, aggr, prev, bnd, mask, std, mean, out=out)
    ts = []
    for _ in range(iters):
        flush.zero_()
        with ops.profile_launches() as prof:
            ops.node_update_step(net.aggr_mlp, out_map, rec, aggr, prev, bnd, mask, std, mean, out=out)
        ts.append(prof.rows[0])
us = sorted(r[1] for r in ts)[len(ts) // 2]
print(f"node_out B={B} {ts[0][0]} {us:.1f} us  {ts[0][2] / 1e6:.1f} MB  {ts[0][2] / us / 1e3:.1f} GB/s")

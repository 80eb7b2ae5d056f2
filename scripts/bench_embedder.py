"""Time the grid embedder (narrow inputs 17|17|18|4 -> 64, tc4.cu narrow-input mode) on the MEPS grid, L2 flushed;
NLAM_TC_TIMELINE=1 prints the in-kernel timeline.  usage: python scripts/bench_embedder.py [B] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from neural_lam_b200 import ops
from neural_lam_b200.networks import make_mlp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
G = 268 * 238
torch.manual_seed(0)
mlp = make_mlp([56, 64, 64]).to(dev)
srcs = [torch.randn(B, G, 17, device=dev), torch.randn(B, G, 17, device=dev), torch.randn(B, G, 18, device=dev),
        torch.randn(G, 4, device=dev).unsqueeze(0).expand(B, -1, -1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(2):
        mlp.apply_rows(srcs)
    ts = []
    for _ in range(iters):
        flush.zero_()
        with ops.profile_launches() as prof:
            mlp.apply_rows(srcs)
        ts.append(prof.rows[0])
us = sorted(r[1] for r in ts)[len(ts) // 2]
print(f"embedder B={B} {ts[0][0]} {us:.1f} us  {ts[0][2] / 1e6:.1f} MB  {ts[0][2] / us / 1e3:.1f} GB/s")

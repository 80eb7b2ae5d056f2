"""Run a few eager forecast steps of the bench workload (target for ncu captures of every kernel of a step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
spec, ds, model, fc = bench.build_model(dev, math="auto")
G = model.num_grid_nodes
torch.manual_seed(0)
prev, prev_prev = torch.randn(B, G, bench.D_STATE, device=dev), torch.randn(B, G, bench.D_STATE, device=dev)
forcing = torch.randn(B, G, bench.D_FORCING, device=dev)
bnd = torch.randn(B, G, bench.D_STATE, device=dev)
with torch.no_grad():
    for _ in range(steps):
        out = model.forward_with_boundary(prev, prev_prev, forcing, bnd, fc.boundary_mask)
        prev_prev, prev = prev, out
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))

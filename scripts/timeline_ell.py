import sys, os
os.environ["NLAM_TC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
spec = synthetic.make_graph_spec(268, 238)
ei = spec["m2g_edge_index"]
torch.manual_seed(0)
net = nlb.InteractionNet(ei, 64, update_edges=False, math="tf32").to(dev)
mesh = torch.randn(B, 6561, 64, device=dev)
grid = torch.randn(B, 268 * 238, 64, device=dev)
edge = torch.randn(1, ei.shape[1], 64, device=dev).expand(B, -1, -1)
with torch.no_grad():
    for i in range(3):
        print("--- call", i, file=sys.stderr)
        net(mesh, grid, edge)
torch.cuda.synchronize()
